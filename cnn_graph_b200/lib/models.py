"""``cgcnn`` -- the Chebyshev graph CNN, with the constructor and the string-bound
``filter`` / ``brelu`` / ``pool`` plugin surface of the reference's ``lib/models.py``
(constructor :61-127, ops :161-274, residual topology :282-378).

The reference module does not import in the fork (``base_model`` is undefined,
lib/models.py:20); this one derives from the ``GraphModel`` harness.  Its ``_inference`` is
the upstream topology that the constructor's own architecture printout describes
(lib/models.py:25-42, 92-111): ``filter -> brelu -> pool`` per graph-conv layer, then the
fully connected stack.  The fork's residual network (lib/models.py:352-378) is available
as ``residual_network`` and is what ``lib.graph_conv.GraphConv`` runs.
"""
import numpy as np
import torch

from .. import ops
from . import variables
from .graph_model import GraphModel


class GraphConvOps(object):
    """The graph-conv building blocks shared by ``cgcnn`` and ``GraphConv``; each keeps the
    reference's name, arguments and variable names ('weights', 'bias')."""

    # when True, b1relu is the fork's plain ReLU (bias lines commented out, lib/models.py:229-235);
    # when False it is upstream's relu(x + b[1, 1, F]).
    b1relu_has_bias = True

    # ---- filters ------------------------------------------------------------------
    def chebyshev5(self, x, L, Fout, K):
        """Chebyshev filter, x [N, M, Fin] -> [N, M, Fout] (lib/models.py:192-224)."""
        N, M, Fin = (int(d) for d in x.shape)
        W = self._weight_variable([Fin * K, Fout], regularization=False)
        return ops.cheb_filter(x, W, L, K, lmax=2)

    def chebyshev2(self, x, L, Fout, K):
        """Same forward as chebyshev5; like the reference (basis through tf.py_func,
        lib/models.py:161-190) no gradient flows to ``x`` -- first-layer use only."""
        N, M, Fin = (int(d) for d in x.shape)
        W = self._weight_variable([Fin * K, Fout], regularization=False)
        return ops.cheb_filter(x, W, L, K, lmax=2, grad_x=False)

    def fourier(self, x, L, Fout, K):
        """Spectral filter in the eigenbasis of L (lib/models.py:146-157): one Fout x Fin matrix per frequency,
        ``weights`` [M, Fout, Fin] without L2; K is ignored (the reference overwrites it with M)."""
        N, M, Fin = (int(d) for d in x.shape)
        W = self._weight_variable([M, Fout, Fin], regularization=False)
        return ops.fourier_filter(x, W, L)

    # ---- bias + nonlinearity --------------------------------------------------------
    def b1relu(self, x):
        """Bias and ReLU, one bias per filter (lib/models.py:226-235)."""
        if not self.b1relu_has_bias:
            return ops.bias_act(x, None, 'relu')
        b = self._bias_variable([1, 1, int(x.shape[2])], regularization=False)
        return ops.bias_act(x, b, 'relu')

    def b1tanh(self, x):
        """tanh(x + b[1, 1, F]) (lib/models.py:237-241)."""
        b = self._bias_variable([1, 1, int(x.shape[2])], regularization=False)
        return ops.bias_act(x, b, 'tanh')

    def b2relu(self, x):
        """Bias and ReLU, one bias per vertex per filter (lib/models.py:243-247)."""
        b = self._bias_variable([1, int(x.shape[1]), int(x.shape[2])], regularization=False)
        return ops.bias_act(x, b, 'relu')

    # dense head through the library's tensor-core GEMM (fp32-level accuracy); False = torch.addmm (cuBLAS fp32)
    fc_on_tensor_cores = True

    # brelu -> pool in one pass over the filter output (and one over its gradient) when both are the stock
    # building blocks; values and gradients are those of pool(brelu(x)).  Set to False to run them separately
    # (then nets['conv*/bias_relu'] is recorded as well).
    fuse_brelu_pool = True

    def _fused_brelu_pool(self, x, p):
        """Returns pool(brelu(x), p) through ops.bias_act_pool, or None when this model's brelu / pool are not
        the stock ones (the plugin surface lets users bind their own)."""
        if not self.fuse_brelu_pool or p <= 1:
            return None
        brelu = getattr(self.brelu, '__func__', None)
        pool = getattr(self.pool, '__func__', None)
        stock_act = {GraphConvOps.b1relu: 'relu', GraphConvOps.b2relu: 'relu', GraphConvOps.b1tanh: 'tanh'}
        stock_pool = {GraphConvOps.mpool1: 'max', GraphConvOps.apool1: 'avg'}
        if brelu not in stock_act or pool not in stock_pool:
            return None
        act, kind = stock_act[brelu], stock_pool[pool]
        if not ops.bias_act_pool_supported(act, p, kind):
            return None
        if brelu is GraphConvOps.b1relu and not self.b1relu_has_bias:
            b = None
        elif brelu is GraphConvOps.b2relu:
            b = self._bias_variable([1, int(x.shape[1]), int(x.shape[2])], regularization=False)
        else:
            b = self._bias_variable([1, 1, int(x.shape[2])], regularization=False)
        return ops.bias_act_pool(x, b, act, p, kind)

    # ---- pooling ----------------------------------------------------------------------
    def mpool1(self, x, p):
        """Max pooling of size p over the permuted vertex axis (lib/models.py:249-257)."""
        return ops.pool(x, p, 'max') if p > 1 else x

    def apool1(self, x, p):
        """Average pooling of size p (lib/models.py:259-266); fake vertices count as zeros."""
        return ops.pool(x, p, 'avg') if p > 1 else x

    # ---- dense head -------------------------------------------------------------------
    def fc(self, x, Mout, relu=True):
        """Fully connected layer (lib/models.py:268-274)."""
        N, Min = (int(d) for d in x.shape)
        W = self._weight_variable([Min, Mout], regularization=True)
        b = self._bias_variable([Mout], regularization=True)
        if x.is_meta:
            return x.new_empty((N, Mout))
        if self.fc_on_tensor_cores and x.is_cuda:
            return ops.linear(x, W, b, relu=relu)         # cg_gemm_f32: bf16 hi+mid split, fp32 accumulate
        x = torch.addmm(b, x, W)
        return torch.relu(x) if relu else x

    def activation_function(self, x, activation):
        """lib/models.py:276-280."""
        if activation == 'brelu':
            return self.brelu(x)
        if activation == 'brelu2':
            return self.b2relu(x)
        if activation == 'tanh':
            return ops.bias_act(x, None, 'tanh')
        raise KeyError(activation)

    # ---- fork topology ----------------------------------------------------------------
    def residual_layer(self, x, nfilter, activation, name_scope):
        """filter -> act -> filter -> (+identity when model_name == 'ResGNN') -> act
        (lib/models.py:282-313)."""
        res = self.model_name == 'ResGNN'
        identity = x
        with self.variable_scope(name_scope):
            with self.variable_scope('sublayer0' if res else 'sublayer0nores'):
                x = self.filter(x, self.L[0], nfilter, self.K[0])
                x = self.activation_function(x, activation)
            with self.variable_scope('sublayer1' if res else 'sublayer1nores'):
                x = self.filter(x, self.L[0], nfilter, self.K[0])
                if res:
                    x = x + identity
                x = self.activation_function(x, activation)
        return x

    def residual_network(self, x):
        """conv_init -> nres x residual_layer -> convN (2 channels), all on L[0]
        (lib/models.py:352-378)."""
        with self.variable_scope('conv_init'):
            x = self.filter(x, self.L[0], self.nfilter, self.K[0])
            x = self.activation_function(x, 'brelu')
        for i in range(self.nres_layer_count):
            x = self.residual_layer(x, self.nfilter, 'brelu', 'residual_layer_{0}'.format(i))
        with self.variable_scope('convN'):
            x = self.filter(x, self.L[0], 2, self.K[0])
        return x

    def stacked_inference(self, x):
        """The fork's two-branch form of ``_inference`` for ``_STACK_NUM > 1`` (lib/graph_conv.py:272-303 ==
        lib/models.py:319-350): channels 0..11 and 12..15 of the input each go through their own residual network
        (scopes final_merge/VC_i), relu, and are summed with one [M, 2] weight per branch (final_merge/VC_i/W_i/weights).
        The 12 / 4 channel split is hard-coded in the reference; more than two branches fail there as well."""
        branches = [x[:, :, 0:12], x[:, :, 12:16]]
        total = None
        with self.variable_scope('final_merge'):
            for i in range(self.stack_num):
                with self.variable_scope('VC_{0}'.format(i)):
                    y = ops.bias_act(self.residual_network(branches[i]), None, 'relu')
                    with self.variable_scope('W_{0}'.format(i)):
                        w = self._weight_variable([int(y.shape[1]), int(y.shape[2])])
                if not y.is_meta:                                # shape tracing declares the variables only
                    y = y * w
                total = y if total is None else total + y
        return total

    # ---- constructor checks shared by cgcnn / GraphConv (lib/models.py:72-85) ----------
    def _select_laplacians(self, L, F, K, p):
        assert len(L) >= len(F) == len(K) == len(p)
        assert np.all(np.array(p) >= 1)
        p_log2 = np.where(np.array(p) > 1, np.log2(p), 0)
        assert np.all(np.mod(p_log2, 1) == 0)            # pool sizes are powers of 2
        assert len(L) >= 1 + np.sum(p_log2)              # enough coarsening levels
        kept, j = [], 0
        for pp in p:
            kept.append(L[j])
            j += int(np.log2(pp)) if pp > 1 else 0
        return kept


class _L2LossSum(torch.autograd.Function):
    """sum_i tf.nn.l2_loss(w_i) = sum_i 0.5 * ||w_i||^2 (upstream cgcnn regulariser) as one autograd node: two
    multi-tensor kernels forward, one backward, instead of a pow / sum / mul / add chain per variable."""

    @staticmethod
    def forward(ctx, *ws):
        ctx.save_for_backward(*ws)
        norms = torch._foreach_norm(list(ws), 2)
        return 0.5 * torch.stack(norms).square().sum()

    @staticmethod
    def backward(ctx, g):
        return tuple(torch._foreach_mul(list(ctx.saved_tensors), g))


def l2_loss_sum(ws):
    return _L2LossSum.apply(*ws)


class cgcnn(GraphConvOps, GraphModel):
    joins_deferred_update = True       # _inference joins a deferred data-parallel update before the dense head (dist.py)

    """Graph CNN with Chebyshev filters.

    L: list of graph Laplacians (one per coarsening level); F, K, p: features, polynomial
    orders and pool sizes of the graph-conv layers; M: sizes of the fully connected layers
    (M[-1] = number of classes).  ``filter`` / ``brelu`` / ``pool`` select the building
    blocks by name, as in the reference (lib/models.py:120-122).
    """

    def __init__(self, L, F, K, p, M, _STACK_NUM=1, _nfilter=64, _nres_layer_count=4, filter='chebyshev5',
                 brelu='b1relu', pool='mpool1', num_epochs=20, learning_rate=0.1, decay_rate=0.95, decay_steps=None,
                 momentum=0.9, regularization=0, dropout=0, batch_size=100, eval_frequency=200, dir_name='',
                 C_0=[1], model_name='ResGNN', verbose=False):
        super().__init__()
        assert _STACK_NUM > 0
        self.nfilter, self.nres_layer_count = _nfilter, _nres_layer_count
        self.stack_num, self.model_name = _STACK_NUM, model_name
        M_0 = L[0].shape[0]
        L = self._select_laplacians(L, F, K, p)
        if verbose:
            self._print_architecture(L, F, K, p, M, M_0, brelu)
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
        self.build_graph(M_0, 1)

    def _print_architecture(self, L, F, K, p, M, M_0, brelu):
        print('NN architecture')
        print('  input: M_0 = {}'.format(M_0))
        for i in range(len(p)):
            F_last = F[i - 1] if i > 0 else 1
            print('  layer {0}: cgconv{0}'.format(i + 1))
            print('    representation: M_{0} * F_{1} / p_{1} = {2} * {3} / {4} = {5}'.format(
                i, i + 1, L[i].shape[0], F[i], p[i], L[i].shape[0] * F[i] // p[i]))
            print('    weights: F_{0} * F_{1} * K_{1} = {2} * {3} * {4} = {5}'.format(
                i, i + 1, F_last, F[i], K[i], F_last * F[i] * K[i]))
        for i in range(len(M)):
            print('  layer {}: {}'.format(len(p) + i + 1, 'logits (softmax)' if i == len(M) - 1 else 'fc{}'.format(i + 1)))
            print('    representation: M_{} = {}'.format(len(p) + i + 1, M[i]))

    # input is [N, M] (one feature per vertex), labels are class ids
    def _input_shape(self, node_num, feature_num):
        return (self.batch_size, node_num)

    def _label_dtype(self):
        return torch.int64

    def _make_optimizer(self):
        # upstream cgcnn: momentum SGD (plain SGD when momentum == 0); on the GPU the whole update is one native launch
        params = list(self.store.parameters())
        if params and all(p.is_cuda and p.dtype == torch.float32 for p in params):
            return ops.NativeMomentumSGD(params, lr=self.learning_rate, momentum=self.momentum)
        return torch.optim.SGD(params, lr=self.learning_rate, momentum=self.momentum)

    def _inference(self, x, dropout):
        """x [N, M] -> logits [N, M[-1]]: (filter, brelu, pool) per layer on L[i], flatten,
        fc(+relu, +dropout) ..., fc (upstream topology; lib/models.py:25-42, 79-111, 268-274)."""
        if x.dim() == 2:
            x = x.unsqueeze(2)                                   # N x M x F=1  (lib/models.py:318)
        for i in range(len(self.p)):
            with self.variable_scope('conv{}'.format(i + 1)):
                fused_layer = self._fused_first_layer(x, i)          # declares the same variables in the same scopes
                if fused_layer is not None:
                    x = fused_layer
                    self.nets['conv{}/pooling'.format(i + 1)] = x.detach()
                    continue
                # upstream wraps the three blocks in tf.name_scope('filter' / 'bias_relu' / 'pooling'), which does not
                # prefix variables: they are conv{i}/weights and conv{i}/bias
                x = self.filter(x, self.L[i], self.F[i], self.K[i])
                fused = self._fused_brelu_pool(x, self.p[i])
                if fused is None:
                    x = self.brelu(x)
                    self.nets['conv{}/bias_relu'.format(i + 1)] = x.detach()    # like self.nets[x.name] = x in the fork: values only --
                    # a tensor with its autograd graph would keep the previous step's AccumulateGrad nodes (and the stream they
                    # were created on) alive into the next step, which breaks CUDA-graph capture of the step now and then
                x = fused if fused is not None else self.pool(x, self.p[i])
                self.nets['conv{}/pooling'.format(i + 1)] = x.detach()
        N, Mv, Fv = (int(d) for d in x.shape)
        x = x.reshape(N, Mv * Fv)
        hook = getattr(self, 'grad_hook', None)
        if hook is not None and hasattr(hook, 'join'):
            hook.join()          # dist.DeferredGradAllReducer: the dense weights get their deferred update before this point
        for i, width in enumerate(self.M[:-1]):
            with self.variable_scope('fc{}'.format(i + 1)):
                x = self.fc(x, width)
                self.nets['fc{}'.format(i + 1)] = x.detach()
                if self.is_train and 0 < dropout < 1 and not x.is_meta:
                    x = torch.nn.functional.dropout(x, p=1 - dropout, training=True)
        with self.variable_scope('logits'):
            x = self.fc(x, self.M[-1], relu=False)
        return x

    def _fused_first_layer(self, x, i):
        """filter -> b1relu -> mpool1(4) of a scalar-input layer as one autograd node (ops.first_layer), or None when the
        layer is not that (user-bound filter / brelu / pool, other pool sizes, an input that needs a gradient, ...)."""
        if x.is_meta or not self.fuse_brelu_pool:
            return None
        stock = (getattr(self.filter, '__func__', None) in (GraphConvOps.chebyshev5, GraphConvOps.chebyshev2) and
                 getattr(self.brelu, '__func__', None) is GraphConvOps.b1relu and
                 getattr(self.pool, '__func__', None) is GraphConvOps.mpool1)
        if not stock or self.p[i] != 4 or int(x.shape[2]) != 1:
            return None
        W = self._weight_variable([self.K[i], self.F[i]], regularization=False)           # [Fin*K, Fout], Fin = 1
        b = self._bias_variable([1, 1, self.F[i]], regularization=False) if self.b1relu_has_bias else None
        if not ops.first_layer_supported(x, W, b, self.L[i], self.K[i], 'relu', 4, 'max'):
            return None
        return ops.first_layer(x, W, b, self.L[i], self.K[i])

    def prediction(self, logits):
        return torch.argmax(logits, dim=1)

    def loss(self, logits, labels, regularization):
        """softmax cross-entropy + L2 on the regularised variables (upstream cgcnn loss)."""
        if logits.is_cuda and logits.dim() == 2 and logits.shape[1] <= 4096:
            ce = ops.softmax_xent(logits, labels)         # loss and its gradient in one native launch
        else:
            ce = torch.nn.functional.cross_entropy(logits, labels)
        if regularization and self.regularizers:
            ce = ce + regularization * l2_loss_sum(self.regularizers)
        return ce

    def evaluate(self, data, labels, sess=None):
        predictions, loss = self.predict(data, labels, sess)
        accuracy = 100.0 * float(np.mean(predictions == labels))
        string = 'accuracy: {:.2f} ({:d} / {:d}), loss: {:.2e}'.format(
            accuracy, int(np.sum(predictions == labels)), len(labels), loss)
        return string, accuracy, 0, loss, predictions
