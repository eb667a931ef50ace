"""Row-partitioned Chebyshev recurrence for graphs too large for one GPU's comfort (BASELINE config C5,
SURVEY.md 8(e)(2)): the M rows of L~ and of every X_k are split into `world` contiguous blocks (after a
locality ordering); one recurrence step needs the rows of X_{k-1} referenced by off-block columns -- the halo --
which the ranks exchange point to point (NCCL all-to-all over NVLink on the B200 box, gloo in CPU tests).
The contraction with W is row-local, so the forward needs no other communication.  The backward pass
(`PartitionedFilter.backward`) runs the same exchange pattern on L~^T for the input gradient and sums the weight
gradient over the ranks with one all-reduce.

Every rank holds the whole (host) operator and derives all send / receive lists from it without talking to
anyone: rank r needs the sorted distinct columns of its row block that fall outside the block; rank p sends, to
every r, the part of r's list that lies in p's block.
"""
import ctypes

import numpy as np
import scipy.sparse
import torch
import torch.distributed as dist

from . import _native
from .dist import shard_bounds


class RowPartition:
    """Partition of a square CSR operator over `world` ranks; local data of `rank`."""

    def __init__(self, L, rank, world):
        L = scipy.sparse.csr_matrix(L, dtype=np.float32)
        L.sum_duplicates()
        L.sort_indices()
        self.M = L.shape[0]
        self.rank, self.world = rank, world
        self.bounds = [shard_bounds(self.M, r, world) for r in range(world)]
        self.r0, self.r1 = self.bounds[rank]
        self.nloc = self.r1 - self.r0
        # halo of every rank (needed to know what to send): sorted distinct off-block columns
        halos = []
        for r, (b, e) in enumerate(self.bounds):
            cols = np.unique(L.indices[L.indptr[b]:L.indptr[e]])
            halos.append(cols[(cols < b) | (cols >= e)])
        self.halo = halos[rank]
        self.nhalo = int(self.halo.size)
        owner_edges = np.array([b for b, _ in self.bounds] + [self.M])
        # receive counts per peer (my halo, grouped by owner: it is sorted, so groups are contiguous and ordered)
        own = np.searchsorted(owner_edges, self.halo, side='right') - 1
        self.recv_counts = [int(np.sum(own == p)) for p in range(world)]
        # send lists per peer: the part of peer's halo inside my block, as local row indices
        self.send_idx = [(halos[p][(halos[p] >= self.r0) & (halos[p] < self.r1)] - self.r0).astype(np.int64)
                         for p in range(world)]
        self.send_counts = [int(s.size) for s in self.send_idx]
        # local operator: rows of my block, columns remapped to [0, nloc) for owned and nloc + position for halo;
        # padded to a square (nloc + nhalo) operator whose halo rows are empty
        sub = L[self.r0:self.r1].tocoo()
        col = sub.col.astype(np.int64)
        inside = (col >= self.r0) & (col < self.r1)
        newcol = np.where(inside, col - self.r0, self.nloc + np.searchsorted(self.halo, col))
        n_ext = self.nloc + self.nhalo
        self.n_ext = n_ext
        self.local = scipy.sparse.csr_matrix((sub.data, (sub.row, newcol)), shape=(n_ext, n_ext), dtype=np.float32)
        self.local.sort_indices()


class PeerHalo:
    """Halo exchange through peer memory instead of a collective ("P2P NVLink reads on symmetric-memory
    buffers"): the [K, rows, C] slab buffer of every rank is allocated with torch's symmetric-memory allocator and mapped
    into all peers; per recurrence step the ranks meet at a device-side barrier (signal pads, no host round trip) and
    every rank PULLS its halo rows of X_{k-1} straight out of the owners' slabs with one kernel (cg_halo_pull: NVLink
    loads, 128-byte lines).  With a tile list from `split_tiles` the pull runs on a side stream under the interior tiles of
    the step (cg_cheb_step_tiles); the boundary tiles follow.  NCCL is not involved in the recurrence at all."""

    def __init__(self, part, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem = symm_mem
        self.part, self.device, self.group = part, device, group if group is not None else dist.group.WORLD
        # rows of the shared buffer: the same on every rank (symmetric allocation)
        n = torch.tensor([part.n_ext], dtype=torch.int64, device=device)
        dist.all_reduce(n, op=dist.ReduceOp.MAX, group=group)
        self.rows = int(n.item())
        owner_edges = np.array([b for b, _ in part.bounds] + [part.M])
        own = np.searchsorted(owner_edges, part.halo, side='right') - 1
        starts = np.array([b for b, _ in part.bounds], dtype=np.int64)
        self.src_rank = torch.as_tensor(own.astype(np.int32), device=device)
        self.src_row = torch.as_tensor((part.halo - starts[own]).astype(np.int32), device=device)
        self.side = torch.cuda.Stream(device=device)
        self._bufs = {}

    def buffer(self, K, C):
        """The symmetric [K, rows, C] buffer for this shape (allocated and exchanged once) and its handle."""
        key = (int(K), int(C))
        if key not in self._bufs:
            t = self.symm_mem.empty((K, self.rows, C), dtype=torch.float32, device=self.device)
            hdl = self.symm_mem.rendezvous(t, self.group)
            self._bufs[key] = (t, hdl)
        return self._bufs[key]

    def barrier(self, hdl):
        hdl.barrier(channel=0)          # device side, on the current stream

    def pull(self, ext, hdl, k, stream=None):
        """ext[k, nloc:nloc + nhalo] <- the owners' rows of slab k."""
        part = self.part
        if part.nhalo == 0:
            return
        K, rows, C = ext.shape
        s = ctypes.c_void_p((stream or torch.cuda.current_stream()).cuda_stream)
        _native.check(_native.lib().cg_halo_pull(ctypes.c_void_p(hdl.buffer_ptrs_dev), self.src_rank.data_ptr(), self.src_row.data_ptr(),
                                                 k * rows * C, ext[k, part.nloc:].data_ptr(), part.nhalo, C, s), 'cg_halo_pull')


def split_tiles(part, tile_rows, device):
    """(interior, boundary) tile lists (int32 device tensors) of this rank's row block: a tile of `tile_rows` rows is
    boundary when one of its rows has an entry in a halo column."""
    ntiles = -(-part.nloc // tile_rows)
    L = part.local
    rows_with_halo = np.unique(np.repeat(np.arange(L.shape[0]), np.diff(L.indptr))[L.indices >= part.nloc])
    bt = np.unique(rows_with_halo[rows_with_halo < part.nloc] // tile_rows)
    mask = np.ones(ntiles, bool)
    mask[bt] = False
    interior = np.nonzero(mask)[0].astype(np.int32)
    return (torch.as_tensor(interior, device=device), torch.as_tensor(bt.astype(np.int32), device=device))


def _exchange(part, x_ext, send_index_dev, group=None):
    """Fill x_ext[nloc:] with the halo rows (x_ext[:nloc] holds this rank's rows); all-to-all of packed rows."""
    C = x_ext.shape[1]
    send = x_ext.index_select(0, send_index_dev) if send_index_dev.numel() else x_ext.new_empty((0, C))
    recv = x_ext[part.nloc:]
    if part.world == 1:
        return
    dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=part.recv_counts,
                           input_split_sizes=part.send_counts, group=group)


class PartitionedBasis:
    """T_k(L~) X for the rows of this rank.  `step_fn(x1_ext, x0_loc_or_None, alpha) -> out_loc` applies the local
    operator; the default is the native CUDA step (cg_cheb_step), tests inject a host function."""

    def __init__(self, L_rescaled, rank=None, world=None, device=None, step_fn=None, group=None, exchange='collective'):
        # rank / world are those of `group` (None = the default process group): the halo exchange runs on it
        # exchange: 'collective' (all_to_all_single of packed rows) or 'peer' (PeerHalo: symmetric memory + pull kernel)
        self.group = group
        if rank is None:
            rank = dist.get_rank(group) if dist.is_initialized() else 0
        if world is None:
            world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.part = RowPartition(L_rescaled, rank, world)
        self.device = device if device is not None else torch.device('cuda', torch.cuda.current_device())
        self.send_index = torch.as_tensor(np.concatenate(self.part.send_idx) if world > 0 else np.zeros(0, np.int64),
                                          dtype=torch.int64, device=self.device)
        self._step_fn = step_fn
        self._handle = None
        if step_fn is None:
            from . import ops
            self._handle = ops.GraphHandle(self.part.local)
        self.peer = None
        self._tiles = {}
        if exchange == 'peer' and self.part.world > 1 and step_fn is None:
            self.peer = PeerHalo(self.part, self.device, group)

    def _native_step(self, x1_ext, x0, alpha, out):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _native.check(_native.lib().cg_cheb_step(self._handle.handle, 0, x1_ext.data_ptr(),
                                                 None if x0 is None else x0.data_ptr(), out.data_ptr(), self.part.nloc,
                                                 x1_ext.shape[1], ctypes.c_float(alpha), stream), 'cg_cheb_step')

    def basis(self, x_loc, K):
        """x_loc [nloc, C] (this rank's rows of X) -> [K, nloc, C] (a view of the [K, nloc + nhalo, C] buffer whose
        tail rows hold the halo of every X_k: each step writes its block in place, no staging copies)."""
        return self.basis_ext(x_loc, K)[:, :self.part.nloc]

    def _native_step_tiles(self, x1_ext, x0, alpha, out, tiles):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _native.check(_native.lib().cg_cheb_step_tiles(self._handle.handle, 0, x1_ext.data_ptr(),
                                                       None if x0 is None else x0.data_ptr(), out.data_ptr(), self.part.nloc,
                                                       x1_ext.shape[1], ctypes.c_float(alpha), tiles.data_ptr(), tiles.numel(),
                                                       stream), 'cg_cheb_step_tiles')

    def _basis_ext_peer(self, x_loc, K):
        """Peer-memory form: barrier, pull on a side stream under the interior tiles, boundary tiles."""
        part, peer = self.part, self.peer
        C = x_loc.shape[1]
        ext, hdl = peer.buffer(K, C)
        if C not in self._tiles:
            tr = _native.lib().cg_cheb_step_tile_rows(self._handle.handle, 0, C)
            self._tiles[C] = split_tiles(part, tr, x_loc.device) if tr > 0 else None
        tiles = self._tiles[C]
        main = torch.cuda.current_stream()
        peer.barrier(hdl)                       # nobody still reads the slabs of the previous call
        ext[0, :part.nloc] = x_loc
        for k in range(1, K):
            peer.barrier(hdl)                   # every rank has written its rows of X_{k-1}
            x0 = ext[k - 2, :part.nloc] if k > 1 else None
            alpha = 2.0 if k > 1 else 1.0
            out = ext[k, :part.nloc]
            if tiles is None:
                peer.pull(ext, hdl, k - 1)
                self._native_step(ext[k - 1], x0, alpha, out)
            else:
                peer.side.wait_stream(main)
                with torch.cuda.stream(peer.side):
                    peer.pull(ext, hdl, k - 1, peer.side)
                self._native_step_tiles(ext[k - 1], x0, alpha, out, tiles[0])       # interior: no halo row needed
                main.wait_stream(peer.side)
                self._native_step_tiles(ext[k - 1], x0, alpha, out, tiles[1])       # boundary
        return ext

    def basis_ext(self, x_loc, K):
        """The whole [K, nloc + nhalo, C] buffer (rows [0, nloc) of every slab are this rank's rows)."""
        if self.peer is not None:
            return self._basis_ext_peer(x_loc, K)
        part = self.part
        C = x_loc.shape[1]
        ext = torch.empty((K, part.n_ext, C), dtype=torch.float32, device=x_loc.device)
        ext[0, :part.nloc] = x_loc
        for k in range(1, K):
            _exchange(part, ext[k - 1], self.send_index, group=self.group)
            x0 = ext[k - 2, :part.nloc] if k > 1 else None
            alpha = 2.0 if k > 1 else 1.0
            if self._step_fn is not None:
                ext[k, :part.nloc] = self._step_fn(ext[k - 1], x0, alpha)
            else:
                self._native_step(ext[k - 1], x0, alpha, ext[k, :part.nloc])
        return ext


class PartitionedFilter:
    """Chebyshev filter of one signal on a row-partitioned graph (config C5: N = 1, x [M, Fin], W [Fin*K, Fout] with
    row fin*K + k as in lib/models.py:222), forward and backward, for the rows of this rank:

        forward    X_k = T_k(L~) x (halo exchange per step)         y_loc  = sum_k X_k[loc] W_k
        backward   Z_k = T_k(L~^T) gy (same pattern on L~^T)        dx_loc = sum_k Z_k[loc] W_k^T
                   dW  = sum over ranks of sum_k X_k[loc]^T gy_loc  (one all-reduce)

    L~ is symmetric only to rounding and not at all for directed graphs, so the backward uses its own partition
    of the true transpose.  `step_fn` / `contract_fn` / `dw_fn` replace the native CUDA
    pieces (cg_cheb_step, cg_cheb_contract, cg_cheb_contract_dw) in host-side tests."""

    def __init__(self, L_rescaled, K, rank=None, world=None, device=None, step_fn=None, step_fn_t=None,
                 contract_fn=None, dw_fn=None, group=None, exchange='collective'):
        L_rescaled = scipy.sparse.csr_matrix(L_rescaled, dtype=np.float32)
        self.K = int(K)
        self.group = group
        self.fwd = PartitionedBasis(L_rescaled, rank, world, device, step_fn, group=group, exchange=exchange)
        self.bwd = PartitionedBasis(scipy.sparse.csr_matrix(L_rescaled.T), rank, world, device, step_fn_t, group=group,
                                    exchange=exchange)
        if self.fwd.peer is not None:
            self.exchange_kind = ('peer memory: device barrier + cg_halo_pull over NVLink (symmetric memory), halo pulled under the '
                                  'interior tiles of the step')
        else:
            self.exchange_kind = 'all_to_all_single (%s)' % (dist.get_backend(group) if dist.is_initialized() else 'single rank')
        self.part = self.fwd.part
        self._contract_fn, self._dw_fn = contract_fn, dw_fn
        self._saved = None

    # stack: [K, n_ext, F] buffer whose first nloc rows per slab are this rank's rows
    def _contract(self, ext, W, transposed):
        K, n_ext, F = ext.shape
        nloc = self.part.nloc
        Fin, Fout = (W.shape[0] // K, W.shape[1])
        if self._contract_fn is not None:
            return self._contract_fn(ext[:, :nloc], W, transposed)
        lib = _native.lib()
        out = torch.empty((nloc, Fin if transposed else Fout), dtype=torch.float32, device=ext.device)
        nbytes = lib.cg_cheb_contract_workspace_bytes(nloc, Fin, Fout, K)
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=ext.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _native.check(lib.cg_cheb_contract(ext.data_ptr(), n_ext * F, W.data_ptr(), out.data_ptr(), nloc, Fin, Fout, K,
                                           int(transposed), ws.data_ptr(), nbytes, stream), 'cg_cheb_contract')
        return out

    def _dw(self, ext, gy, Fout):
        K, n_ext, Fin = ext.shape
        nloc = self.part.nloc
        if self._dw_fn is not None:
            return self._dw_fn(ext[:, :nloc], gy)
        lib = _native.lib()
        dW = torch.empty((Fin * K, Fout), dtype=torch.float32, device=ext.device)
        nbytes = lib.cg_cheb_contract_workspace_bytes(nloc, Fin, Fout, K)
        ws = torch.empty((max(nbytes, 1),), dtype=torch.uint8, device=ext.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _native.check(lib.cg_cheb_contract_dw(ext.data_ptr(), n_ext * Fin, gy.data_ptr(), dW.data_ptr(), nloc, Fin, Fout,
                                              K, ws.data_ptr(), nbytes, stream), 'cg_cheb_contract_dw')
        return dW

    def forward(self, x_loc, W):
        """x_loc [nloc, Fin] (this rank's rows), W [Fin*K, Fout] (replicated) -> y_loc [nloc, Fout]."""
        if W.shape[0] != x_loc.shape[1] * self.K:
            raise ValueError('W must be [Fin*K, Fout] = [%d, *], got %r' % (x_loc.shape[1] * self.K, tuple(W.shape)))
        ext = self.fwd.basis_ext(x_loc.contiguous(), self.K)
        self._saved = (ext, W)
        return self._contract(ext, W.contiguous(), False)

    def backward(self, gy_loc, need_dx=True):
        """gy_loc [nloc, Fout] -> (dx_loc [nloc, Fin] or None, dW [Fin*K, Fout] summed over all ranks)."""
        if self._saved is None:
            raise RuntimeError('PartitionedFilter.backward called before forward')
        ext, W = self._saved
        self._saved = None
        gy_loc = gy_loc.contiguous()
        dW = self._dw(ext, gy_loc, W.shape[1])
        if self.part.world > 1:
            dist.all_reduce(dW, op=dist.ReduceOp.SUM, group=self.group)
        dx = None
        if need_dx:
            zext = self.bwd.basis_ext(gy_loc, self.K)
            dx = self._contract(zext, W.contiguous(), True)
        return dx, dW
