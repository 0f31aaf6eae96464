"""TEST INFRASTRUCTURE ONLY -- minimal binding of the reference's camera message framework.

oracle/_ref/libcamera_message_framework.so is the reference's own transport
(lib/camera_message_framework_c.cpp, lib/camera_message_framework.cpp, lib/filelock.cpp) compiled
UNMODIFIED by oracle/Makefile.  The drop-in tests use it to feed GPU modules through the real
POSIX-shm seqlock buffer, the way a capture source (core/capture_source.py:183-234) and
ModuleBase._loop (core/base.py:711-844) do.  The reference's own Python binding
(core/bindings/camera_message_framework.py) cannot be imported on the GPU box (no /root/reference
there, and it needs the external auv_python_helpers), so the nine C entry points
(lib/camera_message_framework_c.cpp:25-102) are declared again here.
"""
import os

import cffi
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libcamera_message_framework.so")

ffi = cffi.FFI()
ffi.cdef("""
extern int SUCCESS;
extern int NO_NEW_FRAME;
extern int FRAMEWORK_DELETED;
typedef struct Block Block;
typedef struct FramePlane { size_t width, height, depth, type_size, offset; char name[32]; } FramePlane;
typedef struct Frame { size_t width, height, depth, type_size; uint64_t acquisition_time; uint64_t uid;
                       void *data; size_t total_size; size_t plane_count; FramePlane planes[4]; } Frame;
Block *create_block(const char *direction, const size_t max_entry_size_bytes);
Block *open_block(const char *direction);
void delete_block(Block *block);
int write_frame(Block *block, uint64_t acquisition_time, size_t width, size_t height, size_t depth,
                size_t type_size, const unsigned char *data);
int read_frame(Block *block, Frame *frame, bool block_thread);
Frame *create_frame();
void delete_frame(Frame *frame);
uint64_t frame_size(Frame *frame);
""")

_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ffi.dlopen(LIB_PATH)
    return _lib


class Writer:
    """What a capture source does: create the direction's block and write frames into it."""

    def __init__(self, direction, max_entry_size_bytes):
        self.direction = direction.encode()
        self.block = lib().create_block(self.direction, max_entry_size_bytes)
        if self.block == ffi.NULL:
            raise RuntimeError("create_block failed")

    def write(self, acquisition_time_ms, image):
        image = np.ascontiguousarray(image)
        h, w = image.shape[:2]
        depth = image.shape[2] if image.ndim == 3 else 1
        return lib().write_frame(self.block, int(acquisition_time_ms), w, h, depth, image.dtype.itemsize,
                                 ffi.cast("const unsigned char *", image.ctypes.data))

    def close(self):
        if self.block is not None:
            lib().delete_block(self.block)
            self.block = None


class Reader:
    """What ModuleManager / ModuleReader do: open the block, poll read_frame, view frame->data."""

    def __init__(self, direction):
        self.block = lib().open_block(direction.encode())
        if self.block == ffi.NULL:
            raise RuntimeError("open_block: direction does not exist")
        self.frame = lib().create_frame()

    def read(self):
        """Returns (status, acquisition_time, uint8 view of the library-owned buffer) -- the view is
        valid until the next read (lib/camera_message_framework.cpp:417-419)."""
        st = lib().read_frame(self.block, self.frame, False)
        if st != lib().SUCCESS:
            return st, None, None
        f = self.frame
        n = f.width * f.height * f.depth * f.type_size
        view = np.frombuffer(ffi.buffer(f.data, n), dtype=np.uint8).reshape(f.height, f.width, f.depth)
        return st, int(f.acquisition_time), view

    def data_pointer(self):
        return int(ffi.cast("uintptr_t", self.frame.data))

    def close(self):
        if self.frame is not None:
            lib().delete_frame(self.frame)
            self.frame = None
