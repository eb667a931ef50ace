"""Measurement harness behind ``bench.py`` (one module per BASELINE.json config family).  Not product code: the
product is ``cnn_graph_b200/``; these modules build synthetic workloads, time them and format the JSON line."""
