"""SharedBuffer: the per-player trajectory ring, same interface as the reference class
(include/freeimpala/data_structures.h:191-307) on top of the C ABI's fi_ring_*.

Differences a caller can see: readBatch returns a device-resident Batch (the gathered
[M, slot_bytes] bytes live in HBM; .to_host() copies them out) instead of
vector<vector<char>>; everything else (blocking, FIFO, wraparound, stale tails of short
writes, draining, return values) follows the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import FiBatch, check

ELEMENT_SIZE = _lib.ELEMENT_SIZE  # data_structures.h:35


def _as_bytes_view(data):
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data)
        return a, a.ctypes.data, a.nbytes
    b = bytes(data) if not isinstance(data, (bytes, bytearray)) else data
    buf = (C.c_char * len(b)).from_buffer_copy(b) if len(b) else (C.c_char * 1)()
    return buf, C.addressof(buf), len(b)


class Batch:
    """A gathered batch in HBM: [num_slots, slot_bytes] bytes, valid until the next readBatch."""

    def __init__(self, raw: FiBatch):
        self.raw = raw

    def __len__(self):
        return int(self.raw.num_slots)

    def empty(self) -> bool:  # learner.h:79 `batch.empty()`
        return self.raw.num_slots == 0

    @property
    def slot_bytes(self) -> int:
        return int(self.raw.slot_bytes)

    @property
    def dev_ptr(self) -> int:
        return int(self.raw.dev_ptr or 0)

    def to_host(self) -> np.ndarray:
        out = np.empty((len(self), self.slot_bytes), np.uint8)
        if len(self):
            check(_lib.load().fi_batch_to_host(C.byref(self.raw), out.ctypes.data, out.nbytes), "fi_batch_to_host")
        return out


class SharedBuffer:
    def __init__(self, entry_size: int, capacity: int, device: int = 0, _handle=None, _owner=None):
        self._lib = _lib.load()
        self._owner = _owner  # keeps the learner alive when the ring belongs to it
        self._owned = _handle is None
        self._h = _handle if _handle is not None else self._lib.fi_ring_create(device, entry_size, capacity)
        if not self._h:
            raise _lib.FiError(_lib.FI_ERR_CUDA, "fi_ring_create", _lib.last_error())

    def close(self):
        if self._owned and self._h:
            self._lib.fi_ring_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    @property
    def slot_bytes(self) -> int:
        return self._lib.fi_ring_slot_bytes(self._h)

    @property
    def capacity(self) -> int:
        return self._lib.fi_ring_capacity(self._h)

    # data_structures.h:219-241
    def write(self, data) -> bool:
        keep, ptr, n = _as_bytes_view(data)
        return bool(self._lib.fi_ring_write(self._h, ptr, n))

    # data_structures.h:244-264
    def try_write(self, data) -> bool:
        keep, ptr, n = _as_bytes_view(data)
        return bool(self._lib.fi_ring_try_write(self._h, ptr, n))

    def write_many(self, array: np.ndarray) -> int:
        """Write every row of a C-contiguous 2-D uint8 array as one entry (one boundary crossing)."""
        a = np.ascontiguousarray(array)
        assert a.ndim == 2
        return self._lib.fi_ring_write_many(self._h, a.ctypes.data, a.shape[0], a.strides[0], a.shape[1] * a.itemsize)

    def reserve(self):
        """Zero-copy producer: (numpy view of the pinned slot, ticket)."""
        ticket = C.c_uint64()
        p = self._lib.fi_ring_reserve(self._h, C.byref(ticket))
        if not p:
            raise _lib.FiError(_lib.FI_ERR_STATE, "fi_ring_reserve", _lib.last_error())
        view = np.ctypeslib.as_array((C.c_uint8 * self.slot_bytes).from_address(p))
        return view, ticket.value

    def reserve_many(self, count: int):
        """Reserve `count` consecutive pinned slots: (list of slot addresses, first ticket)."""
        ticket = C.c_uint64()
        ptrs = (C.c_void_p * count)()
        if self._lib.fi_ring_reserve_many(self._h, count, ptrs, C.byref(ticket)) != count:
            raise _lib.FiError(_lib.FI_ERR_ARG, "fi_ring_reserve_many", "count must be in [1, capacity]")
        return list(ptrs), ticket.value

    def commit_many(self, first_ticket: int, count: int, n: int | None = None) -> bool:
        return bool(self._lib.fi_ring_commit_many(self._h, first_ticket, count, self.slot_bytes if n is None else n))

    def commit(self, ticket: int, n: int | None = None) -> bool:
        return bool(self._lib.fi_ring_commit(self._h, ticket, self.slot_bytes if n is None else n))

    # data_structures.h:267-300
    def readBatch(self, batch_size: int, stream: int | None = None) -> Batch:
        raw = FiBatch()
        rc = self._lib.fi_ring_read_batch(self._h, batch_size, stream, C.byref(raw))
        check(rc, "fi_ring_read_batch")
        return Batch(raw)

    def setDraining(self) -> None:  # :212-216
        self._lib.fi_ring_set_draining(self._h)

    def getFilledCount(self) -> int:  # :303-306
        return self._lib.fi_ring_filled_count(self._h)
