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
        return torch.optim.Adam(self.store.parameters(), lr=self.learning_rate)

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
        return self.store.vars[name].detach().cpu().numpy()

    # ------------------------------------------------------------------ step / loops
    def _to_device(self, a, dtype=torch.float32):
        if not isinstance(a, np.ndarray):
            a = a.toarray()                                      # scipy sparse batches, :150-151
        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self.device, dtype=dtype, non_blocking=True)

    def train_step(self, batch_data, batch_labels):
        """One optimisation step on device tensors; returns the loss tensor (not synchronised)."""
        self.is_train = True
        for group in self.optimizer.param_groups:
            group['lr'] = self._current_lr()
        self.optimizer.zero_grad(set_to_none=True)
        out = self.inference(batch_data, self.dropout)
        loss = self.loss(out, batch_labels, self.regularization)
        loss.backward()
        if getattr(self, 'grad_hook', None) is not None:
            self.grad_hook(self.store.parameters())      # e.g. data-parallel all-reduce of the gradients
        self.optimizer.step()
        self.global_step += 1
        self.is_train = False
        return loss

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
