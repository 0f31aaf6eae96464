"""Frames from the reference's camera message framework straight to the device.

A capture source writes frames into a POSIX shared-memory file `/dev/shm/auv_visiond_<direction>`
(include/camera_message_framework.hpp:27) laid out as `struct Buffer`
(lib/camera_message_framework.cpp:18-54): a header with the published-frame counter `uid`, three
`FrameMetadata` records carrying the sequence words `v_a` / `v_b`, and three payload slots.  The
reference reader copies the newest slot to the heap (`Block::read_frame`, :423-453), ModuleBase copies
it again (core/base.py:762-768) and a GPU module would copy it a third time.  `CmfRing` maps the same
file, pins the mapping once (`bv_host_register`) and hands `bv_ingest_seqlock` the pointers, so a frame
crosses memory ONCE: DMA from the shared pages to the device, validated by the sequence lock after the
copy, retried when the writer lapped it.

The layout below mirrors the reference's private struct for x86-64 glibc (pthread_cond_t 48 bytes,
pthread_mutex_t 40 bytes); `CmfRing` refuses a file whose size is not
`sizeof(Buffer) + BUFFER_CNT * max_entry_size_bytes`, which is what a changed layout would break.
The transport itself stays the reference's (writers, other readers and the GUI are unaffected).
"""
import ctypes
import mmap
import os

import numpy as np
import torch

from ._ffi import ffi, lib, check, BVError

BLOCK_STUB = "/dev/shm/auv_visiond_"          # include/camera_message_framework.hpp:27
BUFFER_CNT = 3                                # :9
MAX_PLANE_CNT = 4                             # :12
PLANE_NAME_MAX_LEN = 32                       # :15


class PlaneMetadata(ctypes.Structure):        # lib/camera_message_framework.cpp:18-25
    _fields_ = [("width", ctypes.c_uint64), ("height", ctypes.c_uint64), ("depth", ctypes.c_uint64),
                ("type_size", ctypes.c_uint64), ("offset", ctypes.c_uint64), ("name", ctypes.c_char * PLANE_NAME_MAX_LEN)]


class FrameMetadata(ctypes.Structure):        # :27-37
    _fields_ = [("v_a", ctypes.c_uint64), ("v_b", ctypes.c_uint64), ("acquisition_time", ctypes.c_uint64),
                ("total_size", ctypes.c_uint64), ("width", ctypes.c_uint64), ("height", ctypes.c_uint64),
                ("depth", ctypes.c_uint64), ("type_size", ctypes.c_uint64), ("plane_count", ctypes.c_uint64),
                ("planes", PlaneMetadata * MAX_PLANE_CNT)]


class BufferHeader(ctypes.Structure):         # :39-54, up to the flexible `data[]` member (alignas(64))
    _fields_ = [("uid", ctypes.c_uint64), ("max_entry_size_bytes", ctypes.c_size_t), ("deleted", ctypes.c_bool),
                ("metadata", FrameMetadata * BUFFER_CNT), ("cond", ctypes.c_byte * 48), ("cond_mutex", ctypes.c_byte * 40)]


DATA_OFFSET = (ctypes.sizeof(BufferHeader) + 63) // 64 * 64
META_OFFSET = BufferHeader.metadata.offset
META_STRIDE = ctypes.sizeof(FrameMetadata)


class CmfRing:
    """Read side of one direction's shared-memory block, pinned for DMA."""

    def __init__(self, direction, pin=True):
        self.path = BLOCK_STUB + direction
        fd = os.open(self.path, os.O_RDWR)                      # the reference maps readers read-write too (:137)
        try:
            size = os.fstat(fd).st_size
            self._mm = mmap.mmap(fd, size, mmap.MAP_SHARED, mmap.PROT_READ | mmap.PROT_WRITE)
        finally:
            os.close(fd)
        self._view = np.frombuffer(self._mm, dtype=np.uint8)
        self.base = self._view.ctypes.data
        self.header = BufferHeader.from_address(self.base)
        self.slot_bytes = int(self.header.max_entry_size_bytes)
        if size != DATA_OFFSET + BUFFER_CNT * self.slot_bytes:
            self.close()
            raise BVError(-1, "%s: %d bytes is not sizeof(Buffer)=%d + %d x %d: the transport's layout differs from the one "
                              "mirrored here" % (self.path, size, DATA_OFFSET, BUFFER_CNT, self.slot_bytes))
        self._pinned = False
        if pin:
            check(lib.bv_host_register(ffi.cast("void *", self.base), size))
            self._pinned = True
        self.ring = ffi.new("bv_seqlock_ring *")
        self.ring.uid = ffi.cast("const volatile uint64_t *", self.base + BufferHeader.uid.offset)
        self.ring.v_begin = ffi.cast("const volatile uint64_t *", self.base + META_OFFSET + FrameMetadata.v_a.offset)
        self.ring.v_end = ffi.cast("const volatile uint64_t *", self.base + META_OFFSET + FrameMetadata.v_b.offset)
        self.ring.meta_stride = META_STRIDE
        self.ring.data = ffi.cast("const uint8_t *", self.base + DATA_OFFSET)
        self.ring.slot_stride = self.slot_bytes
        self.ring.slots = BUFFER_CNT
        self.last_uid = 0
        self.retries = 0

    @property
    def deleted(self):
        return bool(self.header.deleted)

    def published(self):
        return int(self.header.uid)

    def _meta(self, uid):
        m = self.header.metadata[uid % BUFFER_CNT]
        planes = [(int(p.width), int(p.height), int(p.depth), int(p.type_size), int(p.offset), p.name.decode(errors="replace"))
                  for p in m.planes[:int(m.plane_count)]]
        return int(m.v_b), int(m.acquisition_time), planes

    def ingest(self, ctx, out=None, plane=0, swap_rb=False, max_retries=16):
        """The newest frame's `plane` as a device tensor [H,W,3] on ctx's stream.  Returns
        (tensor, acquisition_time, uid) or None when nothing newer than the last ingested frame exists
        (the reference's NO_NEW_FRAME)."""
        for _ in range(max_retries + 1):
            uid = self.published()
            if uid == 0 or uid == self.last_uid:
                return None
            seq, t_acq, planes = self._meta(uid)
            if plane >= len(planes):
                continue                                        # metadata of a slot that is being rewritten
            w, h, c, ts, off, _name = planes[plane]
            if ts != 1 or c not in (3, 4) or h * w * c + off > self.slot_bytes:
                raise BVError(-1, "plane %d is %dx%dx%d with %d-byte elements: not an 8-bit 3/4-channel image" % (plane, h, w, c, ts))
            if out is None or tuple(out.shape) != (h, w, 3):
                out = ctx.empty((h, w, 3))
            uid_out, tries = ffi.new("uint64_t *"), ffi.new("int *")
            check(lib.bv_ingest_seqlock(ctx.handle, self.ring, off, ffi.cast("uint8_t *", out.data_ptr()), h, w, c,
                                        1 if swap_rb else 0, max_retries, uid_out, tries))
            self.retries += int(tries[0])
            # the frame that was copied is uid_out; its metadata must be the one the shape was taken from
            seq2, t2, planes2 = self._meta(int(uid_out[0]))
            if int(uid_out[0]) == uid or (planes2 == planes):
                self.last_uid = int(uid_out[0])
                return out, (t_acq if int(uid_out[0]) == uid else t2), self.last_uid
        raise BVError(-6, "CmfRing.ingest: no consistent frame after %d attempts" % (max_retries + 1))

    def close(self):
        if getattr(self, "_pinned", False):
            lib.bv_host_unregister(ffi.cast("void *", self.base))
            self._pinned = False
        self.header = None
        self.ring = None
        self._view = None
        try:
            self._mm.close()
        except (BufferError, AttributeError, ValueError):
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
