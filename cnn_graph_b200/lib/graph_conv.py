"""``GraphConv`` -- the fork's importable residual graph-conv regression net, with the
constructor of the reference's ``lib/graph_conv.py:14-81`` and its topology
(``_inference`` :269-303, ``residual_network`` :305-330).  ``b1relu`` is the fork's plain
ReLU here (bias lines commented out in lib/graph_conv.py:181-187).
"""
import numpy as np
import torch

from .graph_model import GraphModel
from .models import GraphConvOps


class GraphConv(GraphConvOps, GraphModel):
    b1relu_has_bias = False

    def __init__(self, L, F, K, p, M, _STACK_NUM=1, _nfilter=64, _nres_layer_count=4, filter='chebyshev5',
                 brelu='b1relu', pool='mpool1', num_epochs=20, learning_rate=0.1, decay_rate=0.95, decay_steps=None,
                 momentum=0.9, regularization=0, dropout=0, batch_size=100, eval_frequency=200, dir_name='',
                 C_0=[6], model_name='ResGNN'):
        super().__init__()
        assert _STACK_NUM > 0
        self.nfilter, self.nres_layer_count = _nfilter, _nres_layer_count
        self.stack_num, self.model_name = _STACK_NUM, model_name
        M_0 = L[0].shape[0]
        L = self._select_laplacians(L, F, K, p)
        self.L, self.F, self.K, self.p, self.M = L, F, K, p, M
        self.num_epochs, self.learning_rate = num_epochs, learning_rate
        self.decay_rate, self.decay_steps, self.momentum = decay_rate, decay_steps, momentum
        self.regularization, self.dropout = regularization, dropout
        self.batch_size, self.eval_frequency = batch_size, eval_frequency
        self.dir_name = dir_name
        self.filter = getattr(self, filter)
        self.brelu = getattr(self, brelu)
        self.pool = getattr(self, pool)
        self.C_0 = C_0
        self.build_graph(M_0, np.sum(self.C_0), 2)              # lib/graph_conv.py:81

    def _inference(self, x, dropout):
        """[N, M, sum(C_0)] -> [N, M, 2] (lib/graph_conv.py:269-303): the residual network, or for
        ``_STACK_NUM > 1`` the fork's two-branch merge over input channels 0..11 / 12..15."""
        if self.stack_num == 1:
            return self.residual_network(x)
        return self.stacked_inference(x)
