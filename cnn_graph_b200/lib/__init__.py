"""Host-side mirror of the reference package ``lib`` for the Chebyshev graph-conv hot path:
same module and function names (``graph``, ``coarsening``, ``filter``, ``models``,
``graph_conv``, ``gconv_lstm``), PyTorch tensors instead of TF tensors, native sm_100a
kernels underneath."""
