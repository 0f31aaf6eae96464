"""YOLO input transform: what Ultralytics does to the frame behind `self.model.track(image)`
(modules/yolo.py:112): LetterBox(640, auto=False) -> BGR2RGB -> HWC2CHW -> half -> /255, batched
over cameras in one kernel launch (the reference does not batch, modules/yolo.py:114)."""
import numpy as np

from ._host import ctx_for, to_device, like_input


def yolo_input(images, new_shape=(640, 640), pad=114, half=True):
    """images: list of np.uint8[H,W,3] BGR frames (sizes may differ) or CUDA tensors.
    Returns float16/float32 [B,3,new_h,new_w] (numpy if the inputs were numpy)."""
    if not images:
        raise ValueError("need at least one image")
    ctx = ctx_for(images[0])
    dev = [to_device(ctx, im) for im in images]
    out = ctx.letterbox(dev, out_h=int(new_shape[0]), out_w=int(new_shape[1]), pad=int(pad), half=half)
    return like_input(ctx, images[0], out)
