"""Model-parallel embedding rows over the GPUs of one box (config C4).

The reference has no such mode: its multi-GPU deployment is one full cache per device over one shared host
parameter server (hps_backend/src/model_state.cpp:395-419).  When a table exceeds one GPU's HBM its rows are
partitioned here by ``owner(key) = mix64(key).lo * G >> 32`` and a request is served with ONE exchange step:

  1. bucket this rank's keys by owner                      (route_hist/route_scatter kernels, hpsx_route_keys)
  2. all-to-all-v of the per-peer counts, then of the keys (NCCL over NVLink; 8 B/key)
  3. every owner looks up the keys it received             (same probe+gather / pull kernels, hpsx_session_lookup_*)
  4. all-to-all-v of the rows back                         (4·D B/key)
  5. scatter rows to their original request positions      (scatter_rows kernel, hpsx_scatter_rows)

``mode="p2p"`` replaces steps 1-5 by ONE fused exchange over NVLink peer memory (``hpsx_shard_group_*`` in
include/hpsx.h): keys and request positions are stored straight into the owners' inboxes, each owner's probe+gather
kernel stores the rows straight into the requesters' output buffers, and device-side flags order the steps — no NCCL
call and no host round trip on the data path (torch.distributed only exchanges the CUDA IPC handles once).

Python only sequences the calls: every computation is a kernel or a C-ABI function of libhpsx.so, every
exchange a torch.distributed collective (``nccl`` on GPUs; point-to-point ``gloo`` ops on the CPU path, where the
lookup is the host parameter server and the routing is ``hpsx_owner_batch``).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import hps as H


def _all_to_all_v(torch, dist, send, send_counts, recv_counts, group):
    """recv = concat over peers of the slice each peer addressed to this rank.  Rows are dim-0 slices."""
    world = dist.get_world_size(group)
    recv = send.new_empty((int(sum(recv_counts)),) + tuple(send.shape[1:]))
    if dist.get_backend(group) == "nccl":
        dist.all_to_all_single(recv, send, [int(c) for c in recv_counts], [int(c) for c in send_counts], group=group)
        return recv
    # gloo has no all_to_all: pairwise exchange (the local slice is a copy)
    rank = dist.get_rank(group)
    s_off = np.concatenate([[0], np.cumsum(send_counts)]).astype(np.int64)
    r_off = np.concatenate([[0], np.cumsum(recv_counts)]).astype(np.int64)
    recv[r_off[rank]:r_off[rank + 1]] = send[s_off[rank]:s_off[rank + 1]]
    ops, keep = [], []
    for peer in range(world):
        if peer == rank:
            continue
        if send_counts[peer]:
            chunk = send[s_off[peer]:s_off[peer + 1]].contiguous()
            keep.append(chunk)
            ops.append(dist.P2POp(dist.isend, chunk, dist.get_global_rank(group, peer) if group else peer, group))
        if recv_counts[peer]:
            ops.append(dist.P2POp(dist.irecv, recv[r_off[peer]:r_off[peer + 1]],
                                  dist.get_global_rank(group, peer) if group else peer, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv


class ShardedLookup:
    """One table of one model, rows sharded over the ranks of ``group``.  Every rank must have loaded the shard
    ``owner(key) == rank`` into its own ``HPS`` (e.g. ``load_table_procedural_shard``)."""

    def __init__(self, hps: H.HPS, model: str, table: int, dim: int, device: int, group=None, mode: str = "nccl"):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.group = torch, dist, group
        self.hps, self.model, self.table, self.dim, self.device = hps, model, table, int(dim), int(device)
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        self.session = hps.session(model, device)
        self.on_gpu = device >= 0
        self._last = {}
        self._rows, self._rows_ptr = None, 0
        self.mode = mode
        self.p2p = None
        if mode == "p2p":
            if not self.on_gpu:
                raise ValueError("mode='p2p' needs GPU sessions")
            self.p2p = H.ShardGroup(self.session, table, self.rank, self.world, self.dim)
            mine = torch.from_numpy(self.p2p.handle.copy())
            if dist.get_backend(group) == "nccl":
                mine = mine.cuda(device)
            handles = [torch.empty_like(mine) for _ in range(self.world)]
            dist.all_gather(handles, mine, group=group)
            self.p2p.connect_ipc(np.stack([h.cpu().numpy() for h in handles]))
            dist.barrier(group=group)
        elif mode != "nccl":
            raise ValueError(f"unknown mode {mode!r}")

    def close(self):
        """Collective: no rank may free its arena while a peer can still store into it."""
        if self.p2p is not None:
            self.dist.barrier(group=self.group)
            self.p2p.close()
            self.p2p = None

    def _lookup_p2p(self, keys, out):
        torch = self.torch
        n = keys.numel()
        torch.cuda.current_stream().synchronize()  # the session launches on its own stream
        ptr = self.p2p.lookup_ptr(keys, n)
        self._last = None  # exchange statistics are fetched on demand (`last`)
        if self._rows is None or self._rows_ptr != ptr:
            # the group's output buffer never moves: wrap it once, slice per request
            self._rows = torch.as_tensor(H._DeviceRows(ptr, self.p2p.capacity, self.dim, self.p2p), device=keys.device)
            self._rows_ptr = ptr
        rows = self._rows[:n]
        if out is not None:
            out.copy_(rows)
            return out
        return rows

    @property
    def last(self):
        if self._last is None and self.p2p is not None:
            st = self.p2p.stats()
            self._last = {"sent_keys": int(st["keys_sent_remote"]), "received_keys": int(st["keys_received_remote"]),
                          "send_counts": st["sent"], "recv_counts": st["received"], "misses": int(st["misses"])}
        return self._last or {}

    @last.setter
    def last(self, value):
        self._last = value

    # -- step 1: bucket keys by owner ---------------------------------------------------------------------
    def _route(self, keys):
        torch = self.torch
        n = keys.numel()
        if self.on_gpu:
            routed = torch.empty_like(keys)
            perm = torch.empty(n, dtype=torch.int32, device=keys.device)
            counts_d = torch.empty(self.world, dtype=torch.int32, device=keys.device)
            counts = H.route_keys(self.device, keys, n, self.world, routed, perm, counts_d).astype(np.int64)
            return routed, perm, counts
        k = keys.numpy()
        own = H.owner_batch(k, self.world)
        order = np.argsort(own, kind="stable")
        counts = np.bincount(own, minlength=self.world).astype(np.int64)
        return torch.from_numpy(k[order]), torch.from_numpy(order.astype(np.int32)), counts

    # -- step 3: local lookup of the keys this rank owns ---------------------------------------------------
    def _lookup_owned(self, keys):
        torch = self.torch
        m = keys.numel()
        rows = torch.empty((m, self.dim), dtype=torch.float32, device=keys.device)
        if m == 0:
            return rows
        # Skewed ownership can hand one rank more keys than a request of the session may hold (max_batch_size x
        # maxnum_catfeature): serve them in pieces instead of failing between the two all-to-alls, where the other
        # ranks would be left waiting in the row exchange.
        cap = self._capacity()
        for lo in range(0, m, cap):
            hi = min(m, lo + cap)
            kk, rr = keys[lo:hi], rows[lo:hi]
            k = [None] * self.table + [kk if self.on_gpu else kk.numpy()]
            o = [None] * self.table + [rr if self.on_gpu else rr.numpy()]
            c = [0] * self.table + [hi - lo]
            if self.on_gpu:
                self.session.lookup_device_keys(k, o, c)
            else:
                self.session.lookup(k, o, c)
        return rows

    def _capacity(self) -> int:
        if getattr(self, "_cap", None) is None:
            self._cap = max(1, int(self.hps.request_capacity(self.model, self.table)))
        return self._cap

    def lookup(self, keys, out: Optional["object"] = None):
        """keys: int64 tensor [n] of THIS rank's request (CUDA tensor for a GPU session, CPU tensor otherwise).
        Returns/fills out [n, dim] in request order."""
        torch, dist = self.torch, self.dist
        keys = keys.contiguous().view(-1)
        n = keys.numel()
        if self.p2p is not None:
            return self._lookup_p2p(keys, out)
        routed, perm, send_counts = self._route(keys)
        # step 2a: counts
        sc = torch.from_numpy(send_counts.copy())
        sc_dev = sc.to(keys.device)
        rc_dev = torch.empty_like(sc_dev)
        if dist.get_backend(self.group) == "nccl":
            dist.all_to_all_single(rc_dev, sc_dev, group=self.group)
        else:
            gathered = [torch.empty_like(sc_dev) for _ in range(self.world)]
            dist.all_gather(gathered, sc_dev, group=self.group)
            rc_dev = torch.stack([g[self.rank] for g in gathered])
        recv_counts = rc_dev.cpu().numpy().astype(np.int64)
        # step 2b: keys
        recv_keys = _all_to_all_v(torch, dist, routed, send_counts, recv_counts, self.group)
        if self.on_gpu:
            torch.cuda.current_stream().synchronize()  # the session launches on its own stream
        # step 3
        rows = self._lookup_owned(recv_keys)
        # step 4: rows travel the reverse way
        back = _all_to_all_v(torch, dist, rows, recv_counts, send_counts, self.group)
        # step 5
        if out is None:
            out = torch.empty((n, self.dim), dtype=torch.float32, device=keys.device)
        if self.on_gpu:
            H.scatter_rows(self.device, back, perm, n, self.dim, out)
            torch.cuda.current_stream().synchronize()
        else:
            out[perm.long()] = back
        self.last = {"sent_keys": int(n - send_counts[self.rank]), "received_keys": int(recv_counts.sum() - recv_counts[self.rank]),
                     "send_counts": send_counts, "recv_counts": recv_counts}
        return out
