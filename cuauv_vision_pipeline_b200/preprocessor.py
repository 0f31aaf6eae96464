"""GPU mirror of the reference's modules/preprocessor.py::Preprocessor.process (lines 47-151) for
the P0 steps: colour-space posts (51-86), colour balance (87-88), per-channel bias / contrast /
brightness (89-109), elliptical erode / dilate (120-129), resize (136-143).

Gaussian blur (110-114), rotate (130-135) and translate (144-149) run on the device with OpenCV's
8-bit fixed-point arithmetic (csrc/filter.cu).  Gaussian noise (115-119) replays numpy's global
legacy generator (MT19937 + polar method) on the device from numpy.random.get_state() and hands the
advanced state back (csrc/noise.cu), so a seeded run matches the reference frame for frame.

The frame is uploaded once and stays on the device across all enabled steps; the three point
operations (bias, contrast, brightness) are folded on the host into one 256-entry table per
channel and applied in a single pass.
"""
import numpy as np

from . import transform
from ._host import ctx_for, to_device, like_input, on_ctx_stream

DEFAULT_OPTIONS = {
    # name: default           (modules/preprocessor.py:10-41)
    "PPX_grayscale": False, "PPX_lab": False, "PPX_rgb_split": False, "PPX_lab_split": False,
    "PPX_hsv_split": False, "PPX_hls_split": False, "PPX_ycrcb_split": False, "PPX_luv_split": False,
    "PPX_color_correction": False, "PPX_r_bias": 0, "PPX_g_bias": 0, "PPX_b_bias": 0,
    "PPX_contrast": 1, "PPX_brightness": 0, "PPX_gaussian_blur": False, "PPX_gaussian_blur_kernel": 1,
    "PPX_gaussian_noise": 0, "PPX_erode": False, "PPX_erode_kernel": 1, "PPX_dilate": False,
    "PPX_dilate_kernel": 1, "PPX_rotate": 0, "PPX_resize": False, "PPX_resize_width": 512,
    "PPX_resize_height": 512, "PPX_resize_ratio": 1, "PPX_translate_x": 0, "PPX_translate_y": 0,
}


class _Value:
    def __init__(self, value):
        self.value = value


def point_lut(r_bias=0, g_bias=0, b_bias=0, contrast=1, brightness=0):
    """Composes preprocessor.py:89-109 into per-channel tables [3,256] (order B,G,R):
    cv2.add(bias, plane) saturating -> clip(x * contrast).astype(u8) (float64, truncation) ->
    clip(x + float(brightness)).astype(u8)."""
    x = np.arange(256, dtype=np.int64)
    lut = np.empty((3, 256), np.uint8)
    for c, bias in enumerate((b_bias, g_bias, r_bias)):
        v = np.clip(x + int(bias), 0, 255) if bias != 0 else x
        if contrast != 1:
            v = np.clip(v * contrast, 0., 255.).astype(np.uint8).astype(np.int64)
        if brightness != 0:
            v = np.clip(v + float(brightness), 0., 255.).astype(np.uint8).astype(np.int64)
        lut[c] = v.astype(np.uint8)
    return lut


class Preprocessor:
    """`Preprocessor(module)` as in the reference; `module` needs `.post(name, image)` and may
    carry `options_dict` (tuner name -> object with `.value`).  Options can also be set directly:
    `pp.options_dict['PPX_lab'].value = True`."""

    def __init__(self, module=None):
        self.module = module
        self.options_dict = {k: _Value(v) for k, v in DEFAULT_OPTIONS.items()}
        if module is not None:
            if not hasattr(module, "options_dict"):
                module.options_dict = {}
            for k, v in self.options_dict.items():
                module.options_dict[k] = v

    def _opt(self, name):
        return self.options_dict[name].value

    def _post(self, name, ctx, like, t):
        if self.module is not None:
            self.module.post(name, like_input(ctx, like, t))

    def process(self, *images):
        from .color_balance import balance
        out = []
        for mat in images:
            ctx = ctx_for(mat)
            cur = to_device(ctx, mat)
            if self._opt("PPX_rgb_split"):                                   # 51-55
                for k, n in ((2, "r"), (1, "g"), (0, "b")):
                    with on_ctx_stream(ctx):                                 # the slice copy runs behind the upload
                        plane = cur[..., k].contiguous()
                    self._post("PPX_rgb_%s_channel" % n, ctx, mat, plane)
            for flag, code, names in (("PPX_lab_split", "bgr2lab", ("lab_l", "lab_a", "lab_b")),      # 56-60
                                      ("PPX_hsv_split", "bgr2hsv", ("hsv_h", "hsv_s", "hsv_v")),      # 61-65
                                      ("PPX_hls_split", "bgr2hls", ("hls_h", "hls_l", "hls_s")),      # 66-70
                                      ("PPX_ycrcb_split", "bgr2ycrcb", ("ycrcb_y", "ycrcb_cr", "ycrcb_cb"))):  # 71-75
                if self._opt(flag):
                    _, planes = ctx.cvt_color(cur, code, split=True)
                    for n, p in zip(names, planes):
                        self._post("PPX_%s_channel" % n, ctx, mat, p)
            if self._opt("PPX_luv_split"):                                   # 76-80 (<= 1 LSB on 0.004 % of colours)
                _, planes = ctx.cvt_color(cur, "bgr2luv", split=True)
                for n, p in zip(("luv_l", "luv_u", "luv_v"), planes):
                    self._post("PPX_%s_channel" % n, ctx, mat, p)
            if self._opt("PPX_grayscale"):                                   # 81-83
                self._post("PPX_grayscale", ctx, mat, ctx.cvt_color(cur, "bgr2gray"))
            if self._opt("PPX_lab"):                                         # 84-86
                self._post("PPX_lab", ctx, mat, ctx.cvt_color(cur, "bgr2lab"))
            if self._opt("PPX_color_correction"):                            # 87-88
                cur = balance(cur)
            rb, gb, bb = self._opt("PPX_r_bias"), self._opt("PPX_g_bias"), self._opt("PPX_b_bias")
            con, bri = self._opt("PPX_contrast"), self._opt("PPX_brightness")
            if rb != 0 or gb != 0 or bb != 0 or con != 1 or bri != 0:        # 89-109
                cur = ctx.apply_lut(cur, point_lut(rb, gb, bb, con, bri))
            if self._opt("PPX_gaussian_blur"):                               # 110-114
                k = self._opt("PPX_gaussian_blur_kernel") * 2 + 1
                cur = ctx.gaussian_blur(cur, (k, k), 0)
            if self._opt("PPX_gaussian_noise") != 0:                         # 115-119
                cur = ctx.add_gaussian_noise(cur, self._opt("PPX_gaussian_noise"))
            if self._opt("PPX_erode"):                                       # 120-124
                k = self._opt("PPX_erode_kernel") * 2 + 1
                cur = ctx.morph(cur, "erode", transform.elliptic_kernel(k, k))
            if self._opt("PPX_dilate"):                                      # 125-129
                k = self._opt("PPX_dilate_kernel") * 2 + 1
                cur = ctx.morph(cur, "dilate", transform.elliptic_kernel(k, k))
            if self._opt("PPX_rotate") != 0:                                 # 130-135
                cur = transform.rotate(cur, self._opt("PPX_rotate"))
            if self._opt("PPX_resize"):                                      # 136-139
                cur = ctx.resize(cur, self._opt("PPX_resize_width"), self._opt("PPX_resize_height"))
            if self._opt("PPX_resize_ratio") != 1:                           # 140-143
                r = self._opt("PPX_resize_ratio")
                cur = ctx.resize(cur, int(cur.shape[1] * r), int(cur.shape[0] * r))
            if self._opt("PPX_translate_x") != 0 or self._opt("PPX_translate_y") != 0:     # 144-149
                cur = transform.translate(cur, self._opt("PPX_translate_x"), self._opt("PPX_translate_y"))
            out.append(like_input(ctx, mat, cur))
        return out
