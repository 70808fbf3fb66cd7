"""Lexicon-sharded variant lookup (mode 2).

Used only when the index is split across GPUs (e.g. a lexicon too large to replicate).  Every rank
loads the same vocabulary, builds the index for its shard of the anagram keys
(`hash(key) mod world`), scores the WHOLE query batch against that shard, and exports its
per-query survivors.  The exports are exchanged with one NCCL all-gather (over NVLink / NVSwitch),
after which `anl_shard_merge` ranks the union per query with the global max frequency.  Results are
identical to the unsharded model (tests/test_gpu_sharded.py).

torch is plumbing here: device buffers for the exchange and `torch.distributed` for the collective.
"""
import ctypes as C

import torch
import torch.distributed as dist

from . import VariantModel, _capi, _check
from ._capi import lib as _lib


class ShardExport:
    """One shard's export of a scored batch, as torch tensors on the shard's device."""

    def __init__(self, heads, records, gids, flags, n_records, max_per_query):
        self.heads, self.records, self.gids, self.flags = heads, records, gids, flags
        self.n_records, self.max_per_query = n_records, max_per_query


class ShardedVariantModel(VariantModel):
    """VariantModel whose index holds one shard of the lexicon.  Same loading API; `build()` takes the
    shard coordinates, lookups go through score -> export -> all-gather -> merge."""

    def build(self, device=-1, shard=0, n_shards=1):
        _check(_lib().anl_model_build_sharded(self._h, int(device), int(shard), int(n_shards)))
        self.shard, self.n_shards = int(shard), int(n_shards)

    def load_index(self, filename, device=-1):
        """Instead of build(): a shard's index file written by save_index; the shard coordinates come from the file."""
        try:
            super().load_index(filename, device)
        finally:
            s, n = C.c_uint32(), C.c_uint32()
            _lib().anl_model_shard(self._h, C.byref(s), C.byref(n))
            self.shard, self.n_shards = s.value, n.value

    # -- stage 1: score the whole batch against this shard ------------------------------------------
    def score(self, inputs, params, device):
        blob, offs = _capi.pack(list(inputs))
        batch = C.c_void_p()
        _check(_lib().anl_device_batch_create(self._h, blob, _capi.u64ptr(offs), len(inputs), C.byref(params.data), C.byref(batch)))
        _check(_lib().anl_device_batch_run(self._h, batch, None))
        n_rec, mx = C.c_uint64(), C.c_uint32()
        _check(_lib().anl_shard_export_size(self._h, batch, C.byref(n_rec), C.byref(mx)))
        n = len(inputs)
        dev = torch.device("cuda", device)
        heads = torch.empty((n, 2), dtype=torch.int64, device=dev)            # 16 B per query
        records = torch.empty((max(n_rec.value, 1), 2), dtype=torch.int64, device=dev)  # 16 B per record
        gids = torch.empty((max(n_rec.value, 1),), dtype=torch.int32, device=dev)
        flags = torch.empty((n,), dtype=torch.int32, device=dev)
        _check(_lib().anl_shard_export(self._h, batch, heads.data_ptr(), records.data_ptr(), gids.data_ptr(), flags.data_ptr()))
        return batch, ShardExport(heads, records, gids, flags, n_rec.value, mx.value)

    # -- stage 3: merge all shards' exports ------------------------------------------------------------
    def merge(self, batch, n_queries, heads_all, records_all, gids_all, flags_all, record_stride, max_survivors):
        rs = C.c_void_p()
        try:
            _check(_lib().anl_shard_merge(self._h, batch, self.n_shards, heads_all.data_ptr(), records_all.data_ptr(),
                                          gids_all.data_ptr(), flags_all.data_ptr(), int(record_stride), int(max_survivors),
                                          C.byref(rs)))
            offs = _lib().anl_result_set_offsets(rs)
            var = _lib().anl_result_set_variants(rs)
            return [[(var[j].vocab_id, var[j].dist_score, var[j].freq_score) for j in range(offs[i], offs[i + 1])]
                    for i in range(n_queries)]
        finally:
            if rs:
                _lib().anl_result_set_free(rs)
            _lib().anl_device_batch_free(self._h, batch)

    # -- the exchange inside the library (csrc/shard_comm.cu): NCCL communicator over the shards' ranks ----------
    def init_comm(self, group=None):
        """Collective over the ranks that hold the shards: rank 0 draws an NCCL id, torch.distributed carries it to
        the others (plumbing), every rank joins the library's own communicator with its shard coordinates."""
        rank = dist.get_rank(group)
        world = dist.get_world_size(group)
        assert world == self.n_shards and rank == self.shard, "one rank per shard, rank == shard"
        ident = (C.c_uint8 * 128)()
        if rank == 0:
            _check(_lib().anl_shard_comm_id(ident))
        t = torch.tensor(list(ident), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        ident = (C.c_uint8 * 128)(*t.cpu().tolist())
        _check(_lib().anl_shard_comm_init(self._h, ident, rank, world))
        self._comm = True

    def find_variants_raw(self, inputs, params, device=None, group=None):
        """All ranks call this with the same inputs; every rank returns the full, merged result lists.  One call
        into the library: score against the shard, exchange every shard's survivors over NCCL, merge."""
        inputs = list(inputs)
        if not getattr(self, "_comm", False):
            self.init_comm(group)
        blob, offs = _capi.pack(inputs)
        rs = C.c_void_p()
        _check(_lib().anl_shard_find_variants_batch(self._h, blob, _capi.u64ptr(offs), len(inputs), C.byref(params.data),
                                                    C.byref(rs)))
        try:
            o = _lib().anl_result_set_offsets(rs)
            var = _lib().anl_result_set_variants(rs)
            return [[(var[j].vocab_id, var[j].dist_score, var[j].freq_score) for j in range(o[i], o[i + 1])]
                    for i in range(len(inputs))]
        finally:
            _lib().anl_result_set_free(rs)

    # -- the same exchange with torch.distributed collectives (kept as a cross-check of the library's own) --------
    def find_variants_raw_torch(self, inputs, params, device=None, group=None):
        """All ranks call this with the same inputs; every rank returns the full, merged result lists."""
        inputs = list(inputs)
        device = torch.cuda.current_device() if device is None else device
        world = dist.get_world_size(group)
        assert world == self.n_shards, "one rank per shard"
        batch, ex = self.score(inputs, params, device)
        dev = ex.heads.device
        # sizes first (tiny), then one padded all-gather per array
        sizes = torch.tensor([ex.n_records, ex.max_per_query], dtype=torch.int64, device=dev)
        all_sizes = torch.empty((world, 2), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(all_sizes, sizes, group=group)
        stride = max(1, int(all_sizes[:, 0].max().item()))
        max_surv = int(all_sizes[:, 1].sum().item())
        n = len(inputs)
        rec_pad = torch.zeros((stride, 2), dtype=torch.int64, device=dev)
        rec_pad[: ex.n_records] = ex.records[: ex.n_records]
        gid_pad = torch.zeros((stride,), dtype=torch.int32, device=dev)
        gid_pad[: ex.n_records] = ex.gids[: ex.n_records]
        heads_all = torch.empty((world * n, 2), dtype=torch.int64, device=dev)
        recs_all = torch.empty((world * stride, 2), dtype=torch.int64, device=dev)
        gids_all = torch.empty((world * stride,), dtype=torch.int32, device=dev)
        flags_all = torch.empty((world * n,), dtype=torch.int32, device=dev)
        dist.all_gather_into_tensor(heads_all, ex.heads, group=group)
        dist.all_gather_into_tensor(recs_all, rec_pad, group=group)
        dist.all_gather_into_tensor(gids_all, gid_pad, group=group)
        dist.all_gather_into_tensor(flags_all, ex.flags, group=group)
        torch.cuda.synchronize(dev)
        return self.merge(batch, n, heads_all, recs_all, gids_all, flags_all, stride, max_surv)


def merge_exports_locally(models, batches, exports, n_queries):
    """Single-process emulation of the exchange (all shards on one GPU): concatenates the exports the way
    the all-gather would and merges on shard 0.  Used by the 1-GPU test of the merge path."""
    dev = exports[0].heads.device
    stride = max(1, max(e.n_records for e in exports))
    max_surv = sum(e.max_per_query for e in exports)
    recs, gids = [], []
    for e in exports:
        r = torch.zeros((stride, 2), dtype=torch.int64, device=dev)
        r[: e.n_records] = e.records[: e.n_records]
        g = torch.zeros((stride,), dtype=torch.int32, device=dev)
        g[: e.n_records] = e.gids[: e.n_records]
        recs.append(r)
        gids.append(g)
    heads_all = torch.cat([e.heads for e in exports])
    flags_all = torch.cat([e.flags for e in exports])
    recs_all, gids_all = torch.cat(recs), torch.cat(gids)
    torch.cuda.synchronize(dev)
    out = models[0].merge(batches[0], n_queries, heads_all, recs_all, gids_all, flags_all, stride, max_surv)
    for m, b in list(zip(models, batches))[1:]:
        _lib().anl_device_batch_free(m._h, b)
    return out
