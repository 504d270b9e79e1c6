"""Parallel-tempering swap rounds over replicas partitioned across ranks (one process per GPU).

Mirrors what Pigeons does for `octofit_pigeons` (ext/OctofitterPigeonsExt/OctofitterPigeonsExt.jl:76-128):
replicas carry (ℓ_ref, ℓ_target); adjacent rungs of the β ladder swap with the deterministic even-odd scheme;
β indices move, states do not.  Replicas are block-partitioned: rank r owns [r*n_local, (r+1)*n_local).

backend "nccl": libocto_b200's own ncclAllGather of n_replicas x 2 float64 (octo_pt_swap_round), the
        ncclUniqueId travels over the torch.distributed process group (plumbing only).
backend "gloo": torch.distributed.all_gather on CPU tensors + the library's pure-host octo_pt_decide —
        the same decision code, used by the CPU tests of the N>1 path.
backend "local": single process, all replicas local.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi


class ParallelTempering:
    def __init__(self, n_replicas_total, *, rank=0, world=1, seed=0, beta=None, backend="local", model=None, lib=None):
        if n_replicas_total % world:
            raise ValueError("replicas must divide evenly over ranks (block partition)")
        self.R, self.rank, self.world, self.seed = int(n_replicas_total), int(rank), int(world), int(seed)
        self.n_local = self.R // self.world
        self.beta = np.ascontiguousarray(np.linspace(0.0, 1.0, self.R) if beta is None else beta, dtype=np.float64)
        if self.beta.shape != (self.R,):
            raise ValueError("beta must have one entry per replica")
        self.chain_of_replica = np.arange(self.R, dtype=np.int32)
        self.round = 0
        self.backend = backend
        self._lib = lib if lib is not None else _abi.load_library()
        self._model = model
        if backend == "local" and model is not None:        # single rank with a context: device-ordered rounds work too
            rc = self._lib.octo_pt_init(model._h, None, 0, 1, self.n_local, self.seed)
            if rc:
                raise RuntimeError(self._lib.octo_last_error().decode())
        if backend == "nccl":
            if model is None:
                raise ValueError("backend 'nccl' needs a LogDensityModel (its context owns the communicator)")
            import torch.distributed as dist
            uid = (C.c_char * 128)()
            if world > 1:
                if rank == 0:
                    rc = self._lib.octo_pt_unique_id(uid)
                    if rc:
                        raise RuntimeError(self._lib.octo_last_error().decode())
                box = [bytes(uid.raw)]
                dist.broadcast_object_list(box, src=0)
                uid = (C.c_char * 128).from_buffer_copy(box[0])
            rc = self._lib.octo_pt_init(model._h, uid, self.rank, self.world, self.n_local, self.seed)
            if rc:
                raise RuntimeError(self._lib.octo_last_error().decode())

    @property
    def local_slice(self):
        return slice(self.rank * self.n_local, (self.rank + 1) * self.n_local)

    def local_betas(self):
        """β currently held by each local replica."""
        return self.beta[self.chain_of_replica[self.local_slice]]

    def swap_round(self, ll_ref_local, ll_target_local):
        """One collective swap round; returns the 0/1 acceptance per adjacent rung pair."""
        pair = np.ascontiguousarray(np.stack([ll_ref_local, ll_target_local], axis=1), dtype=np.float64)
        if pair.shape != (self.n_local, 2):
            raise ValueError(f"expected {self.n_local} local replicas")
        acc = np.zeros(max(self.R - 1, 1), dtype=np.int32)
        if self.backend == "nccl":
            rc = self._lib.octo_pt_swap_round(self._model._h, pair.ctypes.data, self.beta.ctypes.data,
                                              self.chain_of_replica.ctypes.data, self.round, acc.ctypes.data)
        else:
            if self.backend == "gloo" and self.world > 1:
                import torch
                import torch.distributed as dist
                parts = [torch.empty(self.n_local, 2, dtype=torch.float64) for _ in range(self.world)]
                dist.all_gather(parts, torch.from_numpy(pair))
                allp = np.ascontiguousarray(torch.cat(parts).numpy())
            else:
                allp = pair
            rc = self._lib.octo_pt_decide(allp.ctypes.data, self.beta.ctypes.data, self.chain_of_replica.ctypes.data,
                                          self.R, self.round, self.seed, acc.ctypes.data)
        if rc:
            raise RuntimeError(self._lib.octo_last_error().decode())
        self.round += 1
        return acc[: self.R - 1]

    def swap_round_device(self, d_pair_local, state, stream=0):
        """The same round in stream order (C ABI `octo_pt_swap_round_device`, backends "nccl" / "local" with a model):
        `d_pair_local` is the DEVICE address of this rank's [n_local x 2] (l_ref, l_target); `state` is a dict of device
        addresses {"ladder", "chain_of_rung", "rung_of_chain", "swap_count", "beta_local"} (see `device_swap_state`).
        One all-gather and one decision kernel are enqueued on `stream`; nothing is synchronised."""
        if self._model is None:
            raise ValueError("swap_round_device needs a LogDensityModel (its context owns the communicator)")
        rc = self._lib.octo_pt_swap_round_device(self._model._h, int(d_pair_local), state["ladder"], state["chain_of_rung"],
                                                 state["rung_of_chain"], state["swap_count"], state.get("beta_local"),
                                                 self.round, int(stream) if stream else None)
        if rc:
            raise RuntimeError(self._lib.octo_last_error().decode())
        self.round += 1

    def device_swap_state(self, torch, device):
        """Device-resident rung assignment for `swap_round_device` (replicated on every rank): torch tensors plus the
        address dict the call takes."""
        t = {"ladder": torch.tensor(self.beta, dtype=torch.float64, device=device),
             "chain_of_rung": torch.arange(self.R, dtype=torch.int32, device=device),
             "rung_of_chain": torch.arange(self.R, dtype=torch.int32, device=device),
             "swap_count": torch.zeros(self.R, dtype=torch.float64, device=device),
             "beta_local": torch.tensor(self.beta[self.local_slice], dtype=torch.float64, device=device)}
        return t, {k: v.data_ptr() for k, v in t.items()}

    def close(self):
        if self._model is not None and self._model._h:
            self._lib.octo_pt_finalize(self._model._h)
