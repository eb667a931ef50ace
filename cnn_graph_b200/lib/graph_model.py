"""Training / evaluation harness with the call surface of the reference's
``lib/graph_model.py`` (``GraphModel``), re-hosted on PyTorch.

Scope note: the harness is not a kernel target.  It exists so
that the drop-in models can be driven exactly like the reference's (``fit`` / ``predict`` /
``evaluate`` / ``get_var``); the graph-conv work inside ``inference`` is native CUDA, the
optimiser and the loss are stock PyTorch.  Differences from the reference that matter for
throughput: batches are copied to the device once per step and nothing but the scalar loss
is fetched back (the reference fetches every tensor in ``self.nets`` each step,
lib/graph_model.py:154-163).
"""
import collections
import time

import numpy as np
import torch

from . import variables


class GraphModel(object):
    """Fork semantics: regression on [N, M, C] signals, MSE loss, Adam, relu prediction."""

    def __init__(self):
        self.regularizers = []
        self.nets = {}
        self.store = None
        self.optimizer = None
        self.global_step = 0
        self.is_train = False
        self.output_num = None
        self.loss_average = None

    # ------------------------------------------------------------------ graph build
    @property
    def device(self):
        return torch.device('cuda', torch.cuda.current_device()) if torch.cuda.is_available() else torch.device('cpu')

    def build_graph(self, node_num, feature_num, output_num=None):
        """Declare every variable by tracing ``inference`` on shape-only (meta) tensors
        (reference lib/graph_model.py:37-59 builds the static TF graph here)."""
        self.output_num = output_num
        self.store = variables.VariableStore(device=self.device, seed=2017)
        self.regularizers = []
        x = torch.empty(self._input_shape(node_num, int(feature_num)), device='meta')
        self.inference(x, self.dropout)
        self.optimizer = self._make_optimizer()
        self.global_step = 0

    def _input_shape(self, node_num, feature_num):
        return (self.batch_size, node_num, feature_num)

    def _make_optimizer(self):
        # lib/graph_model.py:293 -- the fork trains with Adam at the (decayed) learning rate
        params = list(self.store.parameters())
        if params and all(q.is_cuda and q.dtype == torch.float32 for q in params):
            from .. import ops                                   # the whole Adam step in one native launch
            return ops.NativeAdam(params, lr=self.learning_rate)
        return torch.optim.Adam(params, lr=self.learning_rate, capturable=torch.cuda.is_available())

    def _current_lr(self):
        # tf.train.exponential_decay(..., staircase=True), lib/graph_model.py:282-284
        if self.decay_rate != 1 and self.decay_steps:
            return self.learning_rate * self.decay_rate ** int(self.global_step // self.decay_steps)
        return self.learning_rate

    # ------------------------------------------------------------------ model pieces
    def inference(self, data, dropout):
        """Logits / regression output for a batch (reference lib/graph_model.py:210-225)."""
        with variables.use_store(self.store):
            return self._inference(data, dropout)

    def prediction(self, x):
        return torch.relu(x)                                     # lib/graph_model.py:241

    def loss(self, logits, labels, regularization):
        return torch.mean((labels - logits) ** 2)                # lib/graph_model.py:255

    def parameters(self):
        return self.store.parameters()

    def variable_scope(self, name):
        return variables.variable_scope(name)

    def _weight_variable(self, shape, regularization=True):
        """truncated-normal(0, 0.1) variable named 'weights' (lib/graph_model.py:326-333)."""
        var = variables.get_variable('weights', shape, variables.truncated_normal_initializer(0, 0.1))
        if regularization and all(var is not r for r in self.regularizers):
            self.regularizers.append(var)
        return var

    def _bias_variable(self, shape, regularization=True):
        """constant 0.1 variable named 'bias' (lib/graph_model.py:335-342)."""
        var = variables.get_variable('bias', shape, variables.constant_initializer(0.1))
        if regularization and all(var is not r for r in self.regularizers):
            self.regularizers.append(var)
        return var

    def get_var(self, name):
        """Value of a variable by scoped name, e.g. 'conv1/weights' (lib/graph_model.py:199-204)."""
        hook = getattr(self, 'grad_hook', None)
        if hook is not None and hasattr(hook, 'flush'):
            hook.flush()
        return self.store.vars[name].detach().cpu().numpy()

    # ------------------------------------------------------------------ step / loops
    def _to_device(self, a, dtype=torch.float32):
        if not isinstance(a, np.ndarray):
            if self.device.type == 'cuda' and dtype == torch.float32:
                from .. import ops                               # scipy sparse batches (:150-151): CSR to the device, expanded there
                return ops.sparse_batch_to_device(a, self.device)
            a = a.toarray()
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self.device, dtype=dtype, non_blocking=True)

    def train_step(self, batch_data, batch_labels):
        """One optimisation step on device tensors; returns the loss tensor (not synchronised)."""
        self.is_train = True
        for group in self.optimizer.param_groups:
            group['lr'] = self._current_lr()
        self.optimizer.zero_grad(set_to_none=True)
        hook = getattr(self, 'grad_hook', None)
        if hook is not None and hasattr(hook, 'begin_step'):
            hook.begin_step()                            # deferred exchange + update of the previous step's large gradients
        out = self.inference(batch_data, self.dropout)
        loss = self.loss(out, batch_labels, self.regularization)
        loss.backward()
        if hook is not None:
            hook(self.store.parameters())                # e.g. data-parallel all-reduce of the gradients
        self.optimizer.step()
        if hook is not None and hasattr(hook, 'set_lr') and not torch.cuda.is_current_stream_capturing():
            hook.set_lr(self.optimizer.param_groups[0]['lr'])
        self.global_step += 1
        self.is_train = False
        # the value only: a loss that still heads its autograd graph keeps the step's AccumulateGrad nodes (and the stream
        # they were created on) alive for as long as the caller holds it -- into the capture of a later step
        return loss.detach()

    # ------------------------------------------------------------------ captured step / pipelined feeding
    def train_step_graphed(self, batch_data, batch_labels):
        """``train_step`` replayed from a CUDA graph (one launch per step instead of ~80): the step is captured on
        first use -- after three eager warm-up steps, which also build the packed operators -- and re-captured when
        the staircase learning rate or the batch shape changes.  The batch is copied into the graph's static
        input buffers; the returned loss tensor is overwritten by the next replay."""
        lr = self._current_lr()
        key = (tuple(batch_data.shape), batch_data.dtype, tuple(batch_labels.shape), batch_labels.dtype, lr)
        hook = getattr(self, 'grad_hook', None)
        if hook is not None and hasattr(hook, 'begin_step') and hook.active() and not hook.valid:
            # a deferred reducer has nothing pending yet (very first step): the captured graph always applies a pending
            # update, so this one step runs eagerly and leaves one behind.  On a side stream: autograd binds the variables'
            # AccumulateGrad nodes to the stream of their first use and keeps them alive (self.nets holds activations);
            # bound to the legacy default stream they would break the capture that follows.
            if getattr(self, '_eager_stream', None) is None:
                self._eager_stream = torch.cuda.Stream(device=self.device)
            cur, side = torch.cuda.current_stream(), self._eager_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                loss = self.train_step(batch_data, batch_labels)
            cur.wait_stream(side)
            return loss
        cap = getattr(self, '_captured', None)
        if cap is None or cap['key'] != key:
            cap = self._capture_step(batch_data, batch_labels, key)
        if batch_data.data_ptr() != cap['x'].data_ptr():
            cap['x'].copy_(batch_data, non_blocking=True)
        if batch_labels.data_ptr() != cap['y'].data_ptr():
            cap['y'].copy_(batch_labels, non_blocking=True)
        cap['graph'].replay()
        if hook is not None and hasattr(hook, 'set_lr'):
            hook.set_lr(lr)                              # the rate of THIS step, for the deferred update inside the next replay
        self.global_step += 1
        return cap['loss']

    def _capture_step(self, batch_data, batch_labels, key):
        from .. import _native
        x, y = batch_data.clone(), batch_labels.clone()
        # warm-up (allocator, packed operators, optimiser state) must not advance the training: snapshot, restore
        params = list(self.store.parameters())
        step0 = self.global_step
        snap_p = [q.detach().clone() for q in params]
        snap_s = [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in self.optimizer.state.get(q, {}).items()}
                  for q in params]
        hook = getattr(self, 'grad_hook', None)
        snap_h = hook.snapshot() if hook is not None and hasattr(hook, 'snapshot') else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                self.train_step(x, y)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            for q, old_p, old_s in zip(params, snap_p, snap_s):
                q.copy_(old_p)
                state = self.optimizer.state.get(q, {})
                for k in list(state.keys()):
                    v = state[k]
                    if torch.is_tensor(v):
                        if k in old_s and old_s[k] is not None:
                            v.copy_(old_s[k])
                        else:
                            v.zero_()           # created by the warm-up: back to its initial value
                    elif k in old_s:
                        state[k] = old_s[k]
        self.global_step = step0
        if snap_h is not None:
            hook.restore(snap_h)
        self.optimizer.zero_grad(set_to_none=True)
        graph = torch.cuda.CUDAGraph()
        n0 = _native.lib().cg_launch_count()
        with torch.cuda.graph(graph):
            loss = self.train_step(x, y)
        self.global_step = step0            # the capture itself runs nothing
        self._captured = {'key': key, 'graph': graph, 'x': x, 'y': y, 'loss': loss,
                          'native_launches': int(_native.lib().cg_launch_count() - n0)}
        return self._captured

    def graphed_native_launches(self):
        """Native (this library's) kernel launches inside one captured step; 0 before the first capture."""
        cap = getattr(self, '_captured', None)
        return cap['native_launches'] if cap else 0

    def pipelined_trainer(self, perm=None, depth=2, use_graph=True):
        """Feeder for training from pinned host batches: see ``PipelinedTrainer``."""
        return PipelinedTrainer(self, perm=perm, depth=depth, use_graph=use_graph)

    def fit(self, train_data, train_labels, val_data, val_labels):
        """Mini-batch training loop (reference lib/graph_model.py:124-197).  Returns
        (validation scores, validation losses, seconds per step)."""
        t_process, t_wall = time.process_time(), time.time()
        accuracies, losses = [], []
        indices = collections.deque()
        num_steps = int(self.num_epochs * train_data.shape[0] / self.batch_size)
        for step in range(1, num_steps + 1):
            if len(indices) < self.batch_size:                   # use every sample once before reuse
                indices.extend(np.random.permutation(train_data.shape[0]))
            idx = [indices.popleft() for _ in range(self.batch_size)]
            loss = self.train_step(self._to_device(train_data[idx, :]),
                                   self._to_device(train_labels[idx], self._label_dtype()))
            value = float(loss)
            self.loss_average = value if self.loss_average is None else 0.9 * self.loss_average + 0.1 * value
            if step % self.eval_frequency == 0 or step == num_steps:
                epoch = step * self.batch_size / train_data.shape[0]
                print('step {} / {} (epoch {:.2f} / {}):'.format(step, num_steps, epoch, self.num_epochs))
                print('  learning_rate = {:.2e}, loss_average = {:.2e}'.format(self._current_lr(), self.loss_average))
                string, score, _f1, vloss, _ = self.evaluate(val_data, val_labels, sess=True)
                accuracies.append(score)
                losses.append(vloss)
                print('  validation {}'.format(string))
                print('  time: {:.0f}s (wall {:.0f}s)'.format(time.process_time() - t_process, time.time() - t_wall))
        t_step = (time.time() - t_wall) / max(num_steps, 1)
        return accuracies, losses, t_step

    def _label_dtype(self):
        return torch.float32

    def predict(self, data, labels=None, sess=None):
        """Batched forward over a data set, last batch zero-padded (lib/graph_model.py:64-94)."""
        size = data.shape[0]
        outs, loss = [], 0.0
        hook = getattr(self, 'grad_hook', None)
        if hook is not None and hasattr(hook, 'flush'):
            hook.flush()                                 # a deferred data-parallel update must land before the weights are read
        with torch.no_grad():
            for begin in range(0, size, self.batch_size):
                end = min(begin + self.batch_size, size)
                chunk = data[begin:end]
                if not isinstance(chunk, np.ndarray):
                    chunk = chunk.toarray()
                batch = np.zeros((self.batch_size,) + chunk.shape[1:], np.float32)
                batch[:end - begin] = chunk
                out = self.inference(self._to_device(batch), 1)
                if labels is not None:
                    lab = np.zeros((self.batch_size,) + labels.shape[1:], labels.dtype)
                    lab[:end - begin] = labels[begin:end]
                    loss += float(self.loss(out, self._to_device(lab, self._label_dtype()), self.regularization))
                outs.append(self.prediction(out)[:end - begin].cpu().numpy())
        predictions = np.concatenate(outs, axis=0) if outs else np.empty((0,))
        if labels is not None:
            return predictions, loss * self.batch_size / size
        return predictions

    def evaluate(self, data, labels, sess=None):
        """(string, mse, 0, loss, predictions) as the fork returns (lib/graph_model.py:96-122)."""
        t_process, t_wall = time.process_time(), time.time()
        predictions, loss = self.predict(data, labels, sess)
        mse = float(np.sum((labels - predictions) ** 2) / predictions.size)
        string = 'mse: {:.5f} ( {:d}), f1 (weighted), loss: {:.2e}'.format(mse, len(labels), loss)
        if sess is None:
            string += '\ntime: {:.0f}s (wall {:.0f}s)'.format(time.process_time() - t_process, time.time() - t_wall)
        return string, mse, 0, loss, predictions


class PipelinedTrainer(object):
    """Training from HOST batches with the copies off the critical path.

    ``submit(x_host, y_host)`` (pinned tensors; ``x_host`` is the raw [N, M0] signal when ``perm`` is given, as
    the reference feeds ``coarsening.perm_data(X, perm)``, lib/coarsening.py:219) enqueues
      copy stream:     H2D of the batch into one of ``depth`` device buffers (waits until the step that last used
                       the buffer has consumed it)
      compute stream:  wait for the copy, ``cg_perm_data`` into the step's input, the training step (CUDA graph
                       replay, or eager), an asynchronous D2H copy of the loss into a pinned slot
    and returns at once; the host blocks only when all ``depth`` slots are in flight.  ``drain()`` waits for
    everything and returns the losses in submission order.  The reference round-trips every step through
    ``feed_dict`` and fetches every tensor back (lib/graph_model.py:142-163)."""

    def __init__(self, model, perm=None, depth=2, use_graph=True):
        self.model, self.depth, self.use_graph = model, int(depth), bool(use_graph)
        self.device = model.device
        self.perm = None if perm is None else torch.as_tensor(np.asarray(perm, dtype=np.int32), device=self.device)
        self.copy_stream = torch.cuda.Stream(device=self.device)
        self.raw, self.lab, self.x, self.y = [None] * self.depth, [None] * self.depth, None, None
        self.copied = [torch.cuda.Event() for _ in range(self.depth)]
        self.consumed = [torch.cuda.Event() for _ in range(self.depth)]
        self.loss_done = [torch.cuda.Event() for _ in range(self.depth)]
        self.loss_host = torch.zeros(self.depth, dtype=torch.float32).pin_memory()
        self.count, self.losses = 0, []
        self.h2d_bytes_per_step = 0

    def _collect(self, slot):
        self.loss_done[slot].synchronize()
        self.losses.append(float(self.loss_host[slot]))

    def submit(self, x_host, y_host):
        from .. import ops
        i, b = self.count, self.count % self.depth
        if i >= self.depth:
            self._collect(b)                      # bounds the run-ahead; frees the pinned loss slot
        if self.raw[b] is None:
            self.raw[b] = torch.empty(x_host.shape, dtype=x_host.dtype, device=self.device)
            self.lab[b] = torch.empty(y_host.shape, dtype=y_host.dtype, device=self.device)
        self.h2d_bytes_per_step = x_host.numel() * x_host.element_size() + y_host.numel() * y_host.element_size()
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if i >= self.depth:
                self.copy_stream.wait_event(self.consumed[b])
            self.raw[b].copy_(x_host, non_blocking=True)
            self.lab[b].copy_(y_host, non_blocking=True)
            self.copied[b].record(self.copy_stream)
        cur.wait_event(self.copied[b])
        if self.perm is not None:
            if self.x is None:
                self.x = torch.empty((x_host.shape[0], self.perm.numel()), dtype=torch.float32, device=self.device)
            ops.perm_data_device(self.raw[b], None, out=self.x, perm_t=self.perm)
        else:
            if self.x is None:
                self.x = torch.empty_like(self.raw[b])
            self.x.copy_(self.raw[b], non_blocking=True)
        if self.y is None:
            self.y = torch.empty_like(self.lab[b])
        self.y.copy_(self.lab[b], non_blocking=True)
        self.consumed[b].record(cur)
        step = self.model.train_step_graphed if self.use_graph else self.model.train_step
        loss = step(self.x, self.y)
        self.loss_host[b:b + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.loss_done[b].record(cur)
        self.count += 1

    def submit_csr(self, indptr_host, indices_host, values_host, n_cols, y_host):
        """The same for a SPARSE batch: pinned CSR arrays (int32 indptr / indices, float32 values) instead of a dense
        array -- the H2D copy carries only the stored entries (about 1 % of the dense bytes for bag-of-words rows) and
        cg_csr_densify expands them into the step's static input buffer.  No perm (sparse data sets have no coarsening
        here: 20news.ipynb cell 1).  The index arrays may differ in length from batch to batch."""
        from .. import ops
        i, b = self.count, self.count % self.depth
        if i >= self.depth:
            self._collect(b)
        nnz, rows = int(indices_host.numel()), int(indptr_host.numel()) - 1
        slot = self.raw[b]
        if slot is None or slot[1].numel() < nnz or slot[0].numel() != rows + 1:
            cap = max(nnz, 1) * 2
            slot = self.raw[b] = (torch.empty(rows + 1, dtype=torch.int32, device=self.device),
                                  torch.empty(cap, dtype=torch.int32, device=self.device),
                                  torch.empty(cap, dtype=torch.float32, device=self.device))
            self.lab[b] = torch.empty(y_host.shape, dtype=y_host.dtype, device=self.device)
        self.h2d_bytes_per_step = 4 * (rows + 1) + 8 * nnz + y_host.numel() * y_host.element_size()
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if i >= self.depth:
                self.copy_stream.wait_event(self.consumed[b])
            slot[0].copy_(indptr_host, non_blocking=True)
            slot[1][:nnz].copy_(indices_host, non_blocking=True)
            slot[2][:nnz].copy_(values_host, non_blocking=True)
            self.lab[b].copy_(y_host, non_blocking=True)
            self.copied[b].record(self.copy_stream)
        cur.wait_event(self.copied[b])
        if self.x is None:
            self.x = torch.empty((rows, int(n_cols)), dtype=torch.float32, device=self.device)
        ops.csr_densify(slot[0], slot[1], slot[2], int(n_cols), out_rows=rows, out=self.x)
        if self.y is None:
            self.y = torch.empty_like(self.lab[b])
        self.y.copy_(self.lab[b], non_blocking=True)
        self.consumed[b].record(cur)
        step = self.model.train_step_graphed if self.use_graph else self.model.train_step
        loss = step(self.x, self.y)
        self.loss_host[b:b + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        self.loss_done[b].record(cur)
        self.count += 1

    def drain(self):
        first = max(0, self.count - self.depth)
        for i in range(first, self.count):
            self._collect(i % self.depth)
        out, self.losses = self.losses, []
        self.count = 0
        return out
