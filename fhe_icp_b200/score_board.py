"""Peer score board: the gather step of the document-sharded search, fused into the dot-product kernel.

The reference collects per-document scores in a Python loop (/root/reference/batch_operations.py:264-284).
Sharded over the GPUs of one box, the natural form is "every rank evaluates its shard, the encrypted
scores are gathered to the client".  Instead of writing the scores to local HBM, compressing them and
handing them to a collective (three passes + NCCL's kernel competing with an HBM-bound dot product),
the dot-product kernel of every rank stores its finished words -- already in the 32-bit wire form --
straight into the client GPU's memory over NVLink (cudaIpc peer mapping) and publishes an arrival
flag from its last CTA.  The client waits for the flags with a one-warp kernel, decrypts the board and
returns a credit per slot.  No collective, no host synchronisation in the data path.

Layout (client allocation): ``arrive[2][world]`` u64 flags in a 4 KiB header, then
``board[2][world][rows_max][M][stride]`` u32.  Every rank also owns a 4 KiB control block
(``credit[2]`` u64, the kernel's CTA counter, a status word) that the client maps.

Flow control: step ``s`` (1, 2, ...) uses slot ``s % 2``; rank r may overwrite the slot only after the
client consumed step ``s - 2``, which it learns from ``credit[s % 2] >= s - 2`` (a bounded one-warp
wait on a side stream, joined to the compute stream by an event, so it costs no bubble).  All waits
time out (status word -> RuntimeError at the next :meth:`check`), they never hang the GPU, and a
kernel only ever spins on a flag that ANOTHER GPU writes: dependencies between streams of the client's
own GPU (its own shard's arrival, its own credit) are CUDA events.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import _native as N

HEADER_BYTES = 4096
CTRL_BYTES = 4096
_CTRL_COUNTER, _CTRL_STATUS, _CTRL_PTRS = 64, 128, 1024   # byte offsets inside a control block
SLOTS = 2


def slot_of(step: int) -> int:
    return step % SLOTS


def credit_needed(step: int) -> int:
    """Step whose consumption must be acknowledged before ``step`` may overwrite its slot (0: none)."""
    return step - SLOTS if step > SLOTS else 0


def board_offsets(world: int, rows_max: int, M: int, stride: int, slot: int, rank: int) -> Tuple[int, int]:
    """Byte offsets, inside the client allocation, of (arrival flag, first board row) of (slot, rank)."""
    arrive = 8 * (slot * world + rank)
    rows = HEADER_BYTES + 4 * ((slot * world + rank) * rows_max) * M * stride
    return arrive, rows


def board_bytes(world: int, rows_max: int, M: int, stride: int) -> int:
    return HEADER_BYTES + 4 * SLOTS * world * rows_max * M * stride


class _RawDeviceMemory:
    """``__cuda_array_interface__`` view of memory this module allocated through the C-ABI."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _view(ptr: int, shape, typestr: str, dtype, dev) -> torch.Tensor:
    t = torch.as_tensor(_RawDeviceMemory(ptr, shape, typestr), device=dev)
    return t.view(dtype) if t.dtype != dtype else t


class PeerScoreBoard:
    """One per rank.  ``model`` is a compiled FHESimilarityModel (it supplies M, stride and the engine
    handle); ``rows_max`` the largest shard.  Collective construction (handle exchange over the process
    group); afterwards no call of this class communicates through torch.distributed."""

    def __init__(self, model, rows_max: int, client_rank: int = 0, timeout_ms: int = 10000):
        if not model.compiled:
            raise RuntimeError("Model not compiled. Call compile() first.")
        c = model.model.fhe_circuit            # no keygen here: pushing needs the evaluator handle only
        if not c.wire32_supported:
            raise ValueError("score board: the 32-bit wire form would raise the decoding failure probability above "
                             "p_error at these parameters")
        self.model, self.circuit = model, c
        self.M = 2 if c.two_outputs else 1
        self.stride = int(c.lwe.stride)
        self.rows_max = int(rows_max)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.client_rank = int(client_rank)
        self.timeout_ms = int(timeout_ms)
        self.ctx = N.context(model.device)
        self.dev = torch.device("cuda", self.ctx.device)
        self.step = 0
        self._lib = N.lib()
        self._peer_ptrs = []          # mappings to close
        self._own = []                # allocations to free
        self._events = {}
        self._released = [0] * SLOTS    # client: last step released per slot
        self.credit_stream = torch.cuda.Stream(device=self.dev, priority=-1)

        # Every rank: control block; client: board.  Setup is collective, so a rank whose allocation or mapping
        # fails (no peer access between two GPUs, out of memory) must not leave the others waiting in a
        # collective: errors are held back until every rank has gone through the same sequence of collectives,
        # then all ranks raise together.
        is_client = self.rank == self.client_rank
        err: Optional[BaseException] = None
        ctrl_handle = board_handle = None
        self.ctrl = self.board_base = 0
        self._slots = None
        self._own_arrive_tab = None
        try:
            self.ctrl, ctrl_handle = self._alloc(CTRL_BYTES)
            if is_client:
                self.board_base, board_handle = self._alloc(board_bytes(self.world, self.rows_max, self.M, self.stride))
        except Exception as e:   # noqa: BLE001 -- re-raised below, after the collectives
            err = e
        ctrl_ptrs = [self.ctrl]
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, (ctrl_handle, board_handle))
            try:
                if err is None and any(h is None or h[0] is None for h in handles):
                    raise RuntimeError("a peer failed to allocate its control block")
                if err is None and is_client:
                    ctrl_ptrs = [self.ctrl if r == self.rank else self._open(handles[r][0]) for r in range(self.world)]
                elif err is None:
                    if handles[self.client_rank][1] is None:
                        raise RuntimeError("the client rank failed to allocate the score board")
                    self.board_base = self._open(handles[self.client_rank][1])
            except Exception as e:   # noqa: BLE001
                err = e
        try:
            if err is None and is_client:
                # device tables of the credit-flag addresses of every rank, one table per slot
                tab = _view(self.ctrl + _CTRL_PTRS, (SLOTS, self.world), "<i8", torch.int64, self.dev)
                tab.copy_(torch.tensor([[p + 8 * s for p in ctrl_ptrs] for s in range(SLOTS)], dtype=torch.int64))
                self._slots = [
                    _view(self.board_base + board_offsets(self.world, self.rows_max, self.M, self.stride, s, 0)[1],
                          (self.world * self.rows_max, self.M, self.stride), "<i4", torch.int32, self.dev)
                    for s in range(SLOTS)]
            torch.cuda.synchronize(self.dev)
        except Exception as e:   # noqa: BLE001
            err = err or e
        if self.world > 1:   # agreement doubles as the barrier: nobody pushes before every mapping exists
            ok = torch.tensor([0 if err is not None else 1], dtype=torch.int32, device=self.dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if int(ok.item()) == 0:
                self._release_local()
                raise RuntimeError(f"peer score board: setup failed on at least one rank (this rank: {err!r})")
        elif err is not None:
            self._release_local()
            raise err

    def _release_local(self) -> None:
        """Undo this rank's share of a failed setup (no collective)."""
        self._slots = None
        for p in self._peer_ptrs:
            self._lib.fhe_b200_peer_close(self.ctx.handle, C.c_void_p(p))
        for p in self._own:
            self._lib.fhe_b200_peer_free(self.ctx.handle, C.c_void_p(p))
        self._peer_ptrs, self._own = [], []

    # ---- allocation helpers
    def _alloc(self, nbytes: int):
        p = C.c_void_p()
        h = (C.c_uint8 * N.IPC_HANDLE_BYTES)()
        N.check(self._lib.fhe_b200_peer_alloc(self.ctx.handle, nbytes, C.byref(p), h))
        self._own.append(p.value)
        return p.value, bytes(h)

    def _open(self, handle: bytes) -> int:
        p = C.c_void_p()
        h = (C.c_uint8 * N.IPC_HANDLE_BYTES).from_buffer_copy(handle)
        N.check(self._lib.fhe_b200_peer_open(self.ctx.handle, h, C.byref(p)))
        self._peer_ptrs.append(p.value)
        return p.value

    def _st(self, stream=None):
        s = stream if stream is not None else torch.cuda.current_stream(self.dev)
        return s, C.c_void_p(s.cuda_stream)

    # ---- server side
    def push(self, ct, rows: Optional[int] = None) -> int:
        """Evaluate this rank's shard (``ct``: expanded tensor or SeededCiphertexts; None for an empty
        shard) on the current stream with the scores pushed to the client; returns the step number."""
        self.step += 1
        step, slot = self.step, slot_of(self.step)
        main, st = self._st()
        need = credit_needed(step)
        if need and self.rank == self.client_rank:
            # the client's own shard: the slot was released by this process, on this GPU -- an event orders it
            # (kernels that spin on a flag are only ever used ACROSS GPUs, never between streams of one GPU)
            if self._released[slot] < need:
                self.step -= 1
                raise RuntimeError(f"score board: step {step} would overwrite slot {slot} before release() of step {need}")
            main.wait_event(self._events[("consumed", slot)])
        elif need:  # bounded wait for the client's credit on a side stream; joins the compute stream by event
            ev = self._events.setdefault(("credit", slot), torch.cuda.Event())
            with torch.cuda.stream(self.credit_stream):
                N.check(self._lib.fhe_b200_peer_wait(self.ctx.handle, C.c_void_p(self.ctrl + 8 * slot), 1, need,
                                                     self.timeout_ms, C.c_void_p(self.ctrl + _CTRL_STATUS),
                                                     C.c_void_p(self.credit_stream.cuda_stream)))
                ev.record(self.credit_stream)
            main.wait_event(ev)
        arrive, rows_off = board_offsets(self.world, self.rows_max, self.M, self.stride, slot, self.rank)
        from .fhe_similarity import SeededCiphertexts
        B = 0 if ct is None else (ct.bodies.shape[0] if isinstance(ct, SeededCiphertexts) else ct.shape[0])
        if rows is not None:
            B = min(B, int(rows))
        if B > self.rows_max:
            raise ValueError(f"shard of {B} rows exceeds the board's rows_max={self.rows_max}")
        if B == 0:   # empty shard: only the arrival flag
            if self._own_arrive_tab is None:
                self._own_arrive_tab = torch.empty(SLOTS, dtype=torch.int64, device=self.dev)
                self._own_arrive_tab.copy_(torch.tensor(
                    [self.board_base + board_offsets(self.world, self.rows_max, self.M, self.stride, s, self.rank)[0]
                     for s in range(SLOTS)], dtype=torch.int64))
            N.check(self._lib.fhe_b200_peer_signal(self.ctx.handle, C.c_void_p(self._own_arrive_tab.data_ptr() + 8 * slot),
                                                   1, step, st))
            self._mark_pushed(main)
            return step
        push = N.Push(self.board_base + rows_off, self.board_base + arrive, step, self.ctrl + _CTRL_COUNTER)
        h = self.circuit.evaluator_handle()     # key-less: what a server rank holds
        if isinstance(ct, SeededCiphertexts):
            N.check(self._lib.fhe_b200_similarity_run_seeded_push(h, C.c_void_p(ct.bodies.data_ptr()), B, ct.enc_seed,
                                                                  ct.ct_base, C.byref(push), st))
        else:
            N.check(self._lib.fhe_b200_similarity_run_push(h, C.c_void_p(ct.data_ptr()), B, C.byref(push), st))
        self._mark_pushed(main)
        return step

    def _mark_pushed(self, main) -> None:
        if self.rank == self.client_rank:   # collect() orders itself behind the client's own shard by this event
            ev = self._events.setdefault("pushed", torch.cuda.Event())
            ev.record(main)

    # ---- client side
    def collect(self, stream=None) -> torch.Tensor:
        """Client: make ``stream`` wait until every rank's scores of the latest step have arrived; returns
        the slot as an int32 tensor [world * rows_max, M, stride] (rank r's rows start at r * rows_max)."""
        assert self.rank == self.client_rank, "collect() is a client-rank call"
        step, slot = self.step, slot_of(self.step)
        s, st = self._st(stream)
        if "pushed" in self._events:
            s.wait_event(self._events["pushed"])     # own shard: stream order; the kernel below waits for the peers
        N.check(self._lib.fhe_b200_peer_wait(self.ctx.handle, C.c_void_p(self.board_base + 8 * slot * self.world),
                                             self.world, step, self.timeout_ms, C.c_void_p(self.ctrl + _CTRL_STATUS), st))
        return self._slots[slot]

    def release(self, stream=None) -> None:
        """Client: after the decrypt kernels of the latest step are enqueued on ``stream``, hand its slot
        back to every rank (stream-ordered release store of the step into their credit flags)."""
        assert self.rank == self.client_rank, "release() is a client-rank call"
        slot = slot_of(self.step)
        s, st = self._st(stream)
        N.check(self._lib.fhe_b200_peer_signal(self.ctx.handle, C.c_void_p(self.ctrl + _CTRL_PTRS + 8 * slot * self.world),
                                               self.world, self.step, st))
        self._events.setdefault(("consumed", slot), torch.cuda.Event()).record(s)
        self._released[slot] = self.step

    def check(self) -> None:
        """Host check of the time-out status word (synchronises the device)."""
        torch.cuda.synchronize(self.dev)
        status = _view(self.ctrl + _CTRL_STATUS, (1,), "<i4", torch.int32, self.dev)
        if int(status.item()) != 0:
            raise RuntimeError("score board: a wait on a peer flag timed out (a rank died or fell "
                               f"more than {self.timeout_ms} ms behind)")

    def close(self) -> None:
        if not self._own and not self._peer_ptrs:
            return
        torch.cuda.synchronize(self.dev)
        self._slots = None
        for p in self._peer_ptrs:
            self._lib.fhe_b200_peer_close(self.ctx.handle, C.c_void_p(p))
        self._peer_ptrs = []
        if self.world > 1 and dist.is_initialized():
            dist.barrier()     # nobody frees memory a peer still maps
        for p in self._own:
            self._lib.fhe_b200_peer_free(self.ctx.handle, C.c_void_p(p))
        self._own = []
