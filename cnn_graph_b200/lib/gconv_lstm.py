"""Graph-convolutional LSTM with the call surface of the reference's ``lib/gconv_lstm.py``:
``LSTMStateTuple`` (:15-26), ``GConvLSTMCell`` (:29-221, RNN-cell protocol) and
``GconvModel`` (:224-671).

One cell step in the reference is eight independent ``cheby_conv`` graphs (the Chebyshev
basis of ``inputs`` and of ``h`` is recomputed four times each) plus ~10 elementwise
kernels.  Here a step is ONE filter call on ``[x | h]`` against the eight weight matrices
concatenated into ``[(Fin+H)*K, 4H]`` (exactly equivalent because W rows are fin-major:
stacking features stacks rows) followed by ONE gate kernel.  Variable names and shapes are
the reference's (``Wzxt`` ... ``Woht`` ``[K*Fin, H]`` / ``[K*H, H]``, ``bzt`` ... ``bot``).
"""
import collections

import numpy as np
import torch

from .. import ops
from . import filter as filter_module
from . import graph, variables
from .graph_model import GraphModel

_LSTMStateTuple = collections.namedtuple('LSTMStateTuple', ('c', 'h'))


class LSTMStateTuple(_LSTMStateTuple):
    __slots__ = ()

    @property
    def dtype(self):
        c, h = self
        if c.dtype != h.dtype:
            raise TypeError('Inconsistent internal state')
        return c.dtype


class GConvLSTMCell(object):
    """LSTM cell whose eight affine maps are Chebyshev graph filters.

    gate_variant='fork' is the literal lib/gconv_lstm.py:185-215 behaviour (``z = tan(.)``,
    ``o = tanh(.)``); 'standard' is lib/gconvRNN.py:189-213 (``z = tanh``, ``o = sigmoid``).
    ``forget_bias`` is accepted and, as in the reference (:51), never applied.
    """

    def __init__(self, num_units, forget_bias=1.0, state_is_tuple=True, activation=None, reuse=None,
                 laplacian=None, lmax=None, K=None, feat_in=None, nNode=None, filter_type='cheby_conv',
                 gate_variant='fork'):
        self._num_units = num_units
        self._forget_bias = forget_bias
        self._state_is_tuple = state_is_tuple
        self._activation = activation
        self._laplacian = laplacian
        self._lmax = lmax
        self._K = K
        self._feat_in = feat_in
        self._nNode = nNode
        self.filter = getattr(filter_module, filter_type)
        self.gate_variant = gate_variant

    @property
    def state_size(self):
        return (LSTMStateTuple((self._nNode, self._num_units), (self._nNode, self._num_units))
                if self._state_is_tuple else 2 * self._num_units)

    @property
    def output_size(self):
        return self._num_units

    def zero_state(self, batch_size, dtype=torch.float32, device=None):
        if device is None:
            device = torch.device('cuda', torch.cuda.current_device())
        shape = (batch_size, self._nNode, self._num_units)
        return (torch.zeros(shape, dtype=dtype, device=device), torch.zeros(shape, dtype=dtype, device=device))

    def _variables(self, K, feat_in, H):
        uni = variables.random_uniform_initializer(-0.1, 0.1)
        Wx = [variables.get_variable('W%sxt' % g, [K * feat_in, H], uni) for g in 'zifo']
        Wh = [variables.get_variable('W%sht' % g, [K * H, H], uni) for g in 'zifo']
        b = [variables.get_variable('b%st' % g, [H]) for g in 'zifo']
        return Wx, Wh, b

    def __call__(self, inputs, state, scope=None):
        """(inputs [N, M, Fin], (c, h)) -> (new_h, LSTMStateTuple(new_c, new_h))."""
        with variables.variable_scope(scope or type(self).__name__):
            if self._state_is_tuple:
                c, h = state
            else:
                c, h = torch.split(state, state.shape[1] // 2, dim=1)
            K = self._K if self._K is not None else 2
            H = self._num_units
            feat_in = int(inputs.shape[2])
            Wx, Wh, b = self._variables(K, feat_in, H)
            # the eight filters of the reference (lib/gconv_lstm.py:185-207) as two: columns z | i | f | o.
            # x and h are filtered separately and summed, as in the reference; the hidden path then has a
            # feature count (H) that the vectorised tensor-core kernels take, unlike [x | h] (Fin + H)
            bias = torch.cat(b, dim=0)
            pre = (self.filter(inputs, self._laplacian, self._lmax, 4 * H, K, torch.cat(Wx, dim=1)) +
                   self.filter(h, self._laplacian, self._lmax, 4 * H, K, torch.cat(Wh, dim=1)))
            new_h, new_c = ops.lstm_gates(pre, bias, c, self.gate_variant)
            if self._state_is_tuple:
                new_state = LSTMStateTuple(new_c, new_h)
            else:
                new_state = torch.cat([new_c, new_h], dim=1)
            return new_h, new_state


class GconvModel(GraphModel):
    """Human-flow gconv-LSTM regression model (reference lib/gconv_lstm.py:224-671): the
    constructor signature, ``infer_func`` dispatch by name and the layer helpers are kept.
    Dropout between / after LSTM layers is always on with keep probability 0.8, like the
    reference's DropoutWrapper (:616, :623); ``dropout_masks`` lets tests inject the masks.
    """

    def __init__(self, laplacian, seq_num_closeness, seq_num_period, seq_num_trend, filter_num=64, conv_layer_num=4,
                 filter='cheby_conv', num_epochs=20, learning_rate=0.1, decay_rate=0.95, decay_steps=None,
                 momentum=0.9, regularization=0, dropout=0, batch_size=100, eval_frequency=200, dir_name='',
                 feature_num=6, kernel_num=2, in_feature_num=2, out_feature_num=2, infer_func='inference_glstm',
                 lstm_layer_count=1, num_hidden_conv=32, gate_variant='fork', output_keep_prob=0.8):
        super().__init__()
        self.feature_num = feature_num
        self.model_type = 'glstm'
        self.batch_size = batch_size
        self.in_feature_num = in_feature_num
        self.num_time_steps_closeness = seq_num_closeness
        self.num_time_steps_period = seq_num_period
        self.num_time_steps_trend = seq_num_trend
        self.out_feature_num = out_feature_num
        self.laplacian = laplacian
        self.lmax = graph.lmax(self.laplacian)
        self.num_hidden = filter_num
        self.kernel_num = kernel_num
        self.num_epochs, self.learning_rate = num_epochs, learning_rate
        self.decay_rate, self.decay_steps, self.momentum = decay_rate, decay_steps, momentum
        self.regularization, self.dropout = regularization, dropout
        self.eval_frequency = eval_frequency
        self.dir_name = dir_name
        self.filter = filter
        self.node_num = self.laplacian.shape[0]
        self.conv_layer_num = conv_layer_num
        self.infer_func = infer_func
        self.lstm_layer_count = lstm_layer_count
        self.num_hidden_conv = num_hidden_conv
        self.gate_variant = gate_variant
        self.output_keep_prob = output_keep_prob
        self.dropout_masks = None          # optional list (per layer, per step) of [N, M, H] masks
        self.build_graph(self.node_num, int(np.sum(self.feature_num)), self.out_feature_num)

    def to_string(self):
        return '|{0}| {1}| {2}| {3}| {4}| {5}| {6}| {7}| {8}| {9} | {10}| {11}'.format(
            self.feature_num, self.batch_size, self.in_feature_num, self.num_time_steps_closeness, self.num_hidden,
            self.kernel_num, self.learning_rate, self.filter, self.conv_layer_num, self.lstm_layer_count,
            self.infer_func, self.num_hidden_conv)

    def _inference(self, x, dropout):
        return getattr(self, self.infer_func)(x)

    def _filter(self, x, Fout):
        return getattr(filter_module, self.filter)(x, self.laplacian, self.lmax, Fout, self.kernel_num)

    # ---- inference variants ---------------------------------------------------------
    def _unstack_time(self, x):
        N, M, C = (int(d) for d in x.shape)
        T = self.num_time_steps_closeness
        x = x.reshape(N, M, C // T, T)                        # lib/gconv_lstm.py:274, 405
        return [x[..., t] for t in range(T)]

    def inference_glstm(self, x):
        """gLSTM over the closeness frames, then an output filter (lib/gconv_lstm.py:273-283)."""
        outputs = self.glstm_layer(self._unstack_time(x), self.num_time_steps_closeness, self.lstm_layer_count)
        return self.fc_layer(outputs[-1], self.out_feature_num)

    def inference_gconv(self, x):
        """Plain residual graph-conv stack (lib/gconv_lstm.py:298-319)."""
        with self.variable_scope('conv_init'):
            x = ops.bias_act(self._filter(x, self.num_hidden), None, 'relu')
        for i in range(self.conv_layer_num):
            with self.variable_scope('conv_layer_{}'.format(i)):
                x = self.residual_layer(x, self.num_hidden, 'relu', 'residual_layer_{0}'.format(i))
        with self.variable_scope('output_layer'):
            return self._filter(x, self.out_feature_num)

    def inference_glstm_gconv_no_expand(self, x):
        """gLSTM, then conv_init / residual layers / output filter (lib/gconv_lstm.py:402-428)."""
        frames = self._unstack_time(x)
        self.in_feature_num = int(frames[0].shape[2])
        x = self.glstm_layer(frames, self.num_time_steps_closeness, self.lstm_layer_count)[-1]
        with self.variable_scope('conv_init'):
            x = ops.bias_act(self._filter(x, self.num_hidden_conv), None, 'relu')
        for i in range(self.conv_layer_num):
            with self.variable_scope('conv_layer_{}'.format(i)):
                x = self.residual_layer(x, self.num_hidden_conv, 'relu', 'residual_layer_{0}'.format(i))
        with self.variable_scope('output_layer'):
            return self._filter(x, self.out_feature_num)

    # ---- layers -----------------------------------------------------------------------
    def _output_dropout(self, y, layer, step):
        if y.is_meta:
            return y
        keep = self.output_keep_prob
        if self.dropout_masks is not None:
            return y * self.dropout_masks[layer][step]
        if keep >= 1:
            return y
        mask = (torch.rand_like(y) < keep).to(y.dtype) / keep
        return y * mask

    def glstm_layer(self, x, num_time_step, layer_count):
        """Stacked GConvLSTMCells unrolled over the frames (lib/gconv_lstm.py:609-627:
        MultiRNNCell of DropoutWrapper(cell, output_keep_prob=0.8), tf.nn.static_rnn)."""
        cells = []
        for layer in range(layer_count):
            feat_in = self.in_feature_num if layer == 0 else self.num_hidden
            cells.append(GConvLSTMCell(num_units=self.num_hidden, forget_bias=1.0, laplacian=self.laplacian,
                                       lmax=self.lmax, feat_in=feat_in, K=self.kernel_num, nNode=self.node_num,
                                       filter_type=self.filter, gate_variant=self.gate_variant))
        first = x[0]
        N = int(first.shape[0])
        states = []
        for cell in cells:
            shape = (N, self.node_num, self.num_hidden)
            states.append((first.new_zeros(shape), first.new_zeros(shape)))
        outputs = []
        with self.variable_scope('rnn'):
            for step in range(num_time_step):
                y = x[step]
                for layer, cell in enumerate(cells):
                    with self.variable_scope('multi_rnn_cell/cell_{}'.format(layer)):
                        y, new_state = cell(y, states[layer])
                    states[layer] = (new_state.c, new_state.h)
                    y = self._output_dropout(y, layer, step)
                outputs.append(y)
        return outputs

    def fc_layer(self, x, feature_out):
        """Output 'fully connected' layer = one more graph filter (lib/gconv_lstm.py:629-636)."""
        with self.variable_scope('conv_init'):
            return self._filter(x, feature_out)

    def activation_function(self, x, activation):
        return ops.bias_act(x, None, activation)               # lib/gconv_lstm.py:638-640

    def residual_layer(self, x, nfilter, activation, name_scope):
        """lib/gconv_lstm.py:642-671 (the residual branch is always taken there)."""
        identity = x
        with self.variable_scope(name_scope):
            with self.variable_scope('sublayer0'):
                x = self.activation_function(self._filter(x, nfilter), activation)
            with self.variable_scope('sublayer1'):
                x = self._filter(x, nfilter) + identity
                x = self.activation_function(x, activation)
        return x
