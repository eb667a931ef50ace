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
        return (ops.mark_zero(torch.zeros(shape, dtype=dtype, device=device)),
                ops.mark_zero(torch.zeros(shape, dtype=dtype, device=device)))

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
            pre_x = self.filter(inputs, self._laplacian, self._lmax, 4 * H, K, torch.cat(Wx, dim=1))
            # the filter is linear: on the all-zero initial state (zero_state) its output is exactly zero -- skipped;
            # otherwise the two addends are summed inside the gate kernels (no [N, M, 4H] pass for the addition)
            pre_h = None if ops.is_marked_zero(h) else self.filter(h, self._laplacian, self._lmax, 4 * H, K,
                                                                   torch.cat(Wh, dim=1))
            new_h, new_c = ops.lstm_gates(pre_x, bias, c, self.gate_variant, pre2=pre_h)
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
        self.dropout_masks = None          # optional masks [N, M, H]: [layer][step], or flat in DropoutWrapper call order
        self._mask_cursor = 0
        self.build_graph(self.node_num, int(np.sum(self.feature_num)), self.out_feature_num)

    def to_string(self):
        return '|{0}| {1}| {2}| {3}| {4}| {5}| {6}| {7}| {8}| {9} | {10}| {11}'.format(
            self.feature_num, self.batch_size, self.in_feature_num, self.num_time_steps_closeness, self.num_hidden,
            self.kernel_num, self.learning_rate, self.filter, self.conv_layer_num, self.lstm_layer_count,
            self.infer_func, self.num_hidden_conv)

    def _inference(self, x, dropout):
        self._mask_cursor = 0
        return getattr(self, self.infer_func)(x)

    def _filter(self, x, Fout):
        return getattr(filter_module, self.filter)(x, self.laplacian, self.lmax, Fout, self.kernel_num)

    # ---- inference variants (lib/gconv_lstm.py:264-607), selected by name through ``infer_func`` -------------
    def _unstack_time(self, x, T=None, transposed=False):
        """[N, M, C] -> T frames [N, M, C/T]: reshape to [N, M, C/T, T] and unstack the last axis
        (lib/gconv_lstm.py:274-276); ``transposed`` is the extra (0, 1, 3, 2) transpose of :525-527."""
        N, M, C = (int(d) for d in x.shape)
        T = self.num_time_steps_closeness if T is None else T
        x = x.reshape(N, M, C // T, T)
        if transposed:
            x = x.transpose(2, 3)
            return [x[..., t] for t in range(int(x.shape[3]))]
        return [x[..., t] for t in range(T)]

    def _split_periods(self, x, counts):
        """Channel blocks of 2*count features each, in order (closeness | period | trend; :327-331, :469-474)."""
        out, begin = [], 0
        for c in counts:
            out.append(x[:, :, begin:begin + 2 * c])
            begin += 2 * c
        return out

    def _conv_stack(self, x, nfilter, activation, init_activation=None, suffix=''):
        """conv_init -> conv_layer_i residual layers -> output_layer (e.g. :298-319)."""
        with self.variable_scope('conv_init' + suffix):
            x = self.activation_function(self._filter(x, nfilter), init_activation or activation)
        for i in range(self.conv_layer_num):
            with self.variable_scope('conv_layer_{}'.format(i) + suffix):
                x = self.residual_layer(x, nfilter, activation, 'residual_layer_{0}'.format(i))
        with self.variable_scope('output_layer' + suffix):
            return self._filter(x, self.out_feature_num)

    def inference_lstm(self, x, dropout=None):
        return None                                               # lib/gconv_lstm.py:270-271

    def inference_glstm(self, x):
        """gLSTM over the closeness frames, then an output filter (lib/gconv_lstm.py:273-283)."""
        outputs = self.glstm_layer(self._unstack_time(x), self.num_time_steps_closeness, self.lstm_layer_count)
        return self.fc_layer(outputs[-1], self.out_feature_num)

    def inference_glstm_period_no_expand(self, x):
        """lib/gconv_lstm.py:285-296.  As shipped this variant cannot run (float division inside tf.reshape, :288) and
        it returns the LSTM output, not the output filter's; integer division is used here, the return value kept."""
        assert self.num_time_steps_closeness == self.num_time_steps_period
        x = self.glstm_layer(self._unstack_time(x), self.num_time_steps_closeness, self.lstm_layer_count)[-1]
        self.fc_layer(x, self.out_feature_num)
        return x

    def inference_gconv(self, x):
        """Plain residual graph-conv stack with tanh (lib/gconv_lstm.py:298-319)."""
        return self._conv_stack(x, self.num_hidden, 'tanh')

    def inference_gconv_period_no_expand(self, x):
        """Same with relu (lib/gconv_lstm.py:321-342)."""
        return self._conv_stack(x, self.num_hidden, 'relu')

    def inference_gconv_period_expand(self, x):
        """One conv stack per closeness / period / trend block, relu, concat, merge filter (lib/gconv_lstm.py:344-382)."""
        blocks = self._split_periods(x, [self.num_time_steps_closeness, self.num_time_steps_period, self.num_time_steps_trend])
        outs = []
        for j, xb in enumerate(blocks):
            with self.variable_scope('conv_init_{}'.format(j)):
                y = self.activation_function(self._filter(xb, self.num_hidden), 'tanh')
            for i in range(self.conv_layer_num):
                with self.variable_scope('conv_layer_{0}_{1}'.format(i, j)):
                    y = self.residual_layer(y, self.num_hidden, 'relu', 'residual_layer_{0}'.format(i))
            with self.variable_scope('output_layer_{}'.format(j)):
                y = self._filter(y, self.out_feature_num)
            outs.append(self.activation_function(y, 'relu'))
        with self.variable_scope('merge_layer'):
            return self._filter(torch.cat(outs, dim=2), self.out_feature_num)

    def inference_glstm_gconv(self, x):
        """gLSTM, then conv_init / residual layers / output filter at num_hidden_conv (lib/gconv_lstm.py:384-409)."""
        frames = self._unstack_time(x)
        self.in_feature_num = int(frames[0].shape[2])
        x = self.glstm_layer(frames, self.num_time_steps_closeness, self.lstm_layer_count)[-1]
        return self._conv_stack(x, self.num_hidden_conv, 'relu')

    def inference_glstm_gconv_no_expand(self, x):
        """Identical topology (lib/gconv_lstm.py:411-436)."""
        return self.inference_glstm_gconv(x)

    def _merged_lstms(self, x, counts, head=None, transposed=False):
        """merge_i scopes: one gLSTM stack per channel block, last output (optionally through ``head``)."""
        outs = []
        for i, xb in enumerate(self._split_periods(x, counts)):
            with self.variable_scope('merge_{}'.format(i)):
                frames = self._unstack_time(xb, counts[i], transposed)
                y = self.glstm_layer(frames, len(frames), self.lstm_layer_count)[-1]
                outs.append(head(y, i) if head is not None else y)
        return outs

    def inference_glstm_gconv_split(self, x):
        """Two gLSTMs over the first two closeness-sized blocks, concat, conv stack (lib/gconv_lstm.py:439-474)."""
        c = self.num_time_steps_closeness
        outs = self._merged_lstms(x, [c, c])
        return self._conv_stack(torch.cat(outs, dim=2), self.num_hidden, 'relu')

    def _period_counts(self):
        return [self.num_time_steps_closeness, self.num_time_steps_period, self.num_time_steps_trend]

    def _node_weighted_sum(self, outs):
        """X = sum_i x_i * w_i with one [M, Fout] weight per branch (lib/gconv_lstm.py:495-502)."""
        total = None
        for i, y in enumerate(outs):
            with self.variable_scope('merge_{}'.format(i)):
                with self.variable_scope('weight_{}'.format(i)):
                    w = self._weight_variable([int(y.shape[1]), int(y.shape[2])])
            if not y.is_meta:                                # shape tracing declares the variables only
                y = y * w
            total = y if total is None else total + y
        return total

    def inference_glstm_period_expand(self, x):
        """Three gLSTMs, each through the output filter, summed with per-vertex weights (lib/gconv_lstm.py:476-505)."""
        return self._node_weighted_sum(self._merged_lstms(x, self._period_counts(),
                                                          head=lambda y, i: self.fc_layer(y, self.out_feature_num)))

    def inference_glstm_period_expand_gconv1(self, x):
        """Same with the branch filter's weights directly in merge_i (lib/gconv_lstm.py:507-537)."""
        return self._node_weighted_sum(self._merged_lstms(x, self._period_counts(),
                                                          head=lambda y, i: self._filter(y, self.out_feature_num)))

    def inference_glstm_period_expand_gconv2(self, x):
        """Branch frames come from the TRANSPOSED reshape (:525-527), concat, one final filter (lib/gconv_lstm.py:539-566)."""
        outs = self._merged_lstms(x, self._period_counts(), head=lambda y, i: self._filter(y, self.out_feature_num),
                                  transposed=True)
        with self.variable_scope('final'):
            return self._filter(torch.cat(outs, dim=2), self.out_feature_num)

    def inference_glstm_period_expand_gconv3(self, x):
        """Three gLSTMs, concat, conv stack at num_hidden (lib/gconv_lstm.py:568-607)."""
        outs = self._merged_lstms(x, self._period_counts())
        return self._conv_stack(torch.cat(outs, dim=2), self.num_hidden, 'relu')

    # ---- layers -----------------------------------------------------------------------
    def _output_dropout(self, y, layer, step):
        if y.is_meta:
            return y
        keep = self.output_keep_prob
        if self.dropout_masks is not None:
            masks = self.dropout_masks
            if masks and isinstance(masks[0], (list, tuple)):
                return y * masks[layer][step]              # [layer][step] form (one gLSTM stack)
            mask = masks[self._mask_cursor]                # flat form: DropoutWrapper call order (step-major, layer-minor)
            self._mask_cursor += 1
            return y * mask
        if keep >= 1:
            return y
        # Bernoulli(keep) mask scaled by 1 / keep, as tf.nn.dropout: one fused launch forward, one backward
        return torch.nn.functional.dropout(y, p=1.0 - keep, training=True)

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
            states.append((ops.mark_zero(first.new_zeros(shape)), ops.mark_zero(first.new_zeros(shape))))
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
