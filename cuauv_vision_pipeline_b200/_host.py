"""Shared plumbing of the reference-facing mirrors: accept what the reference functions accept
(numpy uint8 arrays, as handed to ModuleBase.process(), core/base.py:762-768) or CUDA tensors,
run on the default context, and hand back the same kind of object.

Stream ordering.  A context runs on its own non-blocking stream (runtime.Context.torch_stream).  A CUDA
tensor handed in by the caller was produced on the caller's CURRENT torch stream, and the caller will
consume a returned tensor there, so both edges are ordered explicitly: on entry the context's stream
waits for the caller's stream, on exit the caller's stream waits for the context's, and the tensors are
recorded on the other stream so that torch's caching allocator does not recycle them early.  All torch
operations of the mirrors themselves (`.contiguous()`, channel slices) run inside `on_ctx_stream(ctx)`.
"""
import contextlib

import numpy as np
import torch

from .runtime import default_context


def is_device(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def ctx_for(x):
    return default_context(x.device.index if is_device(x) else 0)


def on_ctx_stream(ctx):
    """`with on_ctx_stream(ctx):` -- torch work issued inside is enqueued on the context's stream."""
    return torch.cuda.stream(ctx.torch_stream) if ctx is not None else contextlib.nullcontext()


def to_device(ctx, x, dtype=np.uint8):
    if is_device(x):
        caller = torch.cuda.current_stream(x.device)
        if caller != ctx.torch_stream:
            ctx.torch_stream.wait_stream(caller)          # the tensor's producer kernels come first
            x.record_stream(ctx.torch_stream)
        with torch.cuda.stream(ctx.torch_stream):
            return x.contiguous()
    arr = np.asarray(x)
    if dtype == np.uint8 and arr.dtype != np.uint8:
        raise TypeError("expected a uint8 image, got %s" % arr.dtype)
    return ctx.upload(np.ascontiguousarray(arr, dtype=dtype))


def release_to_caller(ctx, t):
    """Make a device result safe to use on the caller's current stream."""
    if is_device(t):
        caller = torch.cuda.current_stream(t.device)
        if caller != ctx.torch_stream:
            caller.wait_stream(ctx.torch_stream)
            t.record_stream(caller)
    return t


def like_input(ctx, x, t):
    """Return `t` (device tensor) as numpy when the caller passed numpy, else ordered for the caller's stream."""
    if is_device(x):
        if isinstance(t, (list, tuple)):
            return type(t)(release_to_caller(ctx, q) for q in t)
        return release_to_caller(ctx, t)
    return ctx.download(t)
