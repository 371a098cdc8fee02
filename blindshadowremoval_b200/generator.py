"""Drop-in ``Generator`` for the reference's Keras model, backed by libbsr.so through ctypes.

Mirrors the call signatures of /root/reference/model.py:228 (``gen(inputs, uv, reg, chuck, training)``)
and /root/reference/model_with_TSM.py:261 (``gen(inputs, uv, reg, frame, share, chuck, training)``) and
returns the same 4-tuple ``(gs, con_rgb, mask22, dif)``.  PyTorch is only a device-memory container:
CUDA tensors go straight to the C ABI as pointers; NumPy arrays take the host path
(``bsr_forward_*_host``: H2D copy, forward, D2H copy).  There is no CPU fallback: if libbsr.so is
missing or no B200 is present this module raises.
"""
from __future__ import annotations

import ctypes
import os
from typing import Dict, Optional

import numpy as np

from . import convert
from .weights import VARIANTS, random_weights

_LIB = None
_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbsr.so")
# 'tc16' = the 16-bit tensor-core product path (binary16 storage, see include/bsr.h); 'bf16' and 'f16' are accepted
# aliases of it ('bf16' is the name round 1 used and north_star's wording), 'fp32check' the CUDA-core check mode.
PRECISIONS = ("tc16", "fp32check")
_PRECISION_ALIASES = {"bf16": "tc16", "f16": "tc16", "fp16": "tc16"}
PLAN_COUNTERS = ("resident", "pinned", "staged", "attn_fused", "graph_replays", "halo3")
IMG = 256
FEAT = 32

_F = ctypes.POINTER(ctypes.c_float)
_SYMBOLS = {
    "bsr_version": (ctypes.c_char_p, []),
    "bsr_act_dtype": (ctypes.c_char_p, []),
    "bsr_convert_h16": (None, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "bsr_check": (ctypes.c_int, [ctypes.c_void_p]),
    "bsr_plan_counter": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "bsr_crc32c": (ctypes.c_uint32, [ctypes.c_uint32, ctypes.c_void_p, ctypes.c_size_t]),
    "bsr_create": (ctypes.c_int, [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_void_p)]),
    "bsr_destroy": (ctypes.c_int, [ctypes.c_void_p]),
    "bsr_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "bsr_load_weights": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t]),
    "bsr_forward_gsc": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 5),
    "bsr_forward_tsm": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 5),
    "bsr_forward_gsc_host": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 4),
    "bsr_forward_tsm_host": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 4),
    "bsr_forward_gsc_host_compact": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 2 + [ctypes.c_int] + [ctypes.c_void_p] * 6),
    "bsr_forward_tsm_host_compact": (ctypes.c_int, [ctypes.c_void_p] + [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 6),
    "bsr_forward_chunk": (ctypes.c_int, [ctypes.c_void_p] * 2 + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 5),
    "bsr_share_layer": (ctypes.c_int, [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_void_p] * 2),
    "bsr_postprocess_ucb": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int] + [ctypes.c_void_p] * 10),
    "bsr_caller_glue": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_int] + [ctypes.c_void_p] * 3),
    "bsr_composite": (ctypes.c_int, [ctypes.c_void_p] * 4 + [ctypes.c_size_t] + [ctypes.c_void_p] * 2),
    "bsr_launch_count": (ctypes.c_int, [ctypes.c_void_p]),
    "bsr_workspace_bytes": (ctypes.c_size_t, [ctypes.c_void_p]),
    "bsr_debug_read": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t)]),
    "bsr_layer_times": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(ctypes.c_float), ctypes.c_int]),
}


def load_library(path: Optional[str] = None):
    """dlopen libbsr.so and bind every symbol of include/bsr.h; raises if anything is missing.
    BSR_LIB=<path> selects another build of the same ABI (e.g. the bfloat16-storage A/B build, `make bf16`)."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = path or os.environ.get("BSR_LIB") or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % path)
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SYMBOLS.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


class BsrError(RuntimeError):
    pass


# tf.split sizes of the dataset chunk by channel count (train_test_GSC.py:419, 806, 870, 900; train_with_TSM.py:675, 717)
CHUNK_LAYOUTS = {
    16: (0, ("img", 3), ("gt", 3), ("uv", 3), ("reg", 6), ("face", 1)),
    17: (1, ("img", 3), ("cmap", 3), ("mask", 1), ("uv", 3), ("reg", 6), ("face", 1)),
    13: (2, ("img", 3), ("uv", 3), ("reg", 6), ("face", 1)),
}


def unpack_chunk(chunk, frames: Optional[int] = None):
    """``tf.reshape(img, [F, 256, 256, -1])`` + ``tf.split(img, [...], 3)`` of the reference's test steps: returns a
    dict of channel-slice views (NumPy or torch) of a [F,256,256,C] chunk, C in {13, 16, 17}; any other shape is first
    reshaped to ``[frames, 256, 256, -1]`` as the reference does (F = 10, or 2 for the SFW mirror pairs)."""
    if chunk.ndim != 4 or tuple(chunk.shape[1:3]) != (IMG, IMG):
        if not frames:
            raise ValueError("a chunk that is not [F,256,256,C] needs `frames`")
        chunk = chunk.reshape(frames, IMG, IMG, -1)
    c = chunk.shape[-1]
    if c not in CHUNK_LAYOUTS:
        raise ValueError("chunk must have 13, 16 or 17 channels, got %d" % c)
    out, o = {}, 0
    for name, width in CHUNK_LAYOUTS[c][1:]:
        out[name] = chunk[..., o:o + width]
        o += width
    return out


def downsample8(x: np.ndarray) -> np.ndarray:
    """``tf.image.resize(x, [32, 32])`` of a [N,256,256,C] map (bilinear, half-pixel centres, no antialias):
    the mean of the centre 2x2 of every 8x8 cell (model.py:237; warp.py:137).  Same operation order as the
    device kernel, so ``forward_compact(u8, downsample8(uv))`` reproduces ``__call__(u8/255, uv)`` bit for bit."""
    x = np.asarray(x, dtype=np.float32)
    if x.ndim != 4 or x.shape[1:3] != (IMG, IMG):
        raise ValueError("expected [N,256,256,C], got %r" % (x.shape,))
    h = np.float32(0.5)
    top = x[:, 3::8, 3::8] * h + x[:, 3::8, 4::8] * h
    bot = x[:, 4::8, 3::8] * h + x[:, 4::8, 4::8] * h
    return np.ascontiguousarray(top * h + bot * h)


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _as_host_array(x):
    """The reference hands the model TF EagerTensors (train_test_GSC.py:419-422); anything that is not a CUDA torch
    tensor but converts to NumPy (``.numpy()`` / ``__array__``) takes the host path."""
    if x is None or isinstance(x, np.ndarray):
        return x
    if getattr(x, "is_cuda", False):
        return x
    if hasattr(x, "numpy"):
        return np.asarray(x.numpy())
    if hasattr(x, "__array__"):
        return np.asarray(x)
    return x


def act_dtype() -> str:
    """Storage type of the 16-bit path of the loaded library: 'f16' (default build) or 'bf16'."""
    return load_library().bsr_act_dtype().decode()


def convert_h16(x: np.ndarray) -> np.ndarray:
    """The library's host-side float -> 16-bit storage rounding (uint16 bit patterns)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.empty(x.shape, np.uint16)
    load_library().bsr_convert_h16(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), x.size)
    return out


class Generator:
    """B200 replacement of ``Generator()`` (train_test_GSC.py:120 / train_with_TSM.py).

    variant   'gsc' (model.py) or 'tsm' (model_with_TSM.py)
    precision 'tc16' (tcgen05 product path; aliases 'bf16', 'f16') or 'fp32check' (CUDA-core check mode)
    """

    def __init__(self, variant: str = "gsc", precision: str = "tc16", device: int = 0, micro_batch: int = 128,
                 weights: Optional[Dict[str, np.ndarray]] = None, seed: Optional[int] = None):
        if variant not in VARIANTS:
            raise ValueError("variant must be one of %r" % (VARIANTS,))
        precision = _PRECISION_ALIASES.get(precision, precision)
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %r" % (PRECISIONS,))
        self.variant, self.precision, self.device = variant, precision, int(device)
        self._lib = load_library()
        h = ctypes.c_void_p()
        rc = self._lib.bsr_create(VARIANTS.index(variant), PRECISIONS.index(precision), self.device, int(micro_batch),
                                  ctypes.byref(h))
        if rc != 0:
            raise BsrError("bsr_create failed (%d): %s" % (rc, self._lib.bsr_last_error(None).decode()))
        self._h = h
        if weights is not None:
            self.load_weights(weights)
        elif seed is not None:
            self.load_weights(random_weights(variant, seed))

    # -- weights ---------------------------------------------------------------------------
    def load_weights(self, weights: Dict[str, np.ndarray]) -> None:
        blob = convert.build_blob(self.variant, weights)
        self.load_blob(blob)

    def load_blob(self, blob: bytes) -> None:
        buf = ctypes.create_string_buffer(blob, len(blob))
        self._check(self._lib.bsr_load_weights(self._h, buf, len(blob)))

    def load_checkpoint(self, index_path: str) -> None:
        """``checkpoint.restore(latest).expect_partial()`` (train_test_GSC.py:362-365) without TF."""
        self.load_blob(convert.convert_checkpoint(index_path, self.variant))

    # -- forward ---------------------------------------------------------------------------
    def __call__(self, inputs, uv, reg=None, frame: Optional[int] = None, share=True, chuck: int = 1,
                 training: bool = False, want=("gs", "con_rgb", "mask22", "dif")):
        if training:
            raise BsrError("training=True is not supported: this is the inference forward only")
        if hasattr(share, "numpy") and not hasattr(share, "item"):
            share = share.numpy()                                   # tf.constant(True) (train_with_TSM.py:676)
        share = bool(share.item() if hasattr(share, "item") else share)
        inputs, uv, reg = _as_host_array(inputs), _as_host_array(uv), _as_host_array(reg)
        host = [isinstance(a, np.ndarray) for a in (inputs, uv, reg) if a is not None]
        if all(host):
            return self._call_host(inputs, uv, reg, frame, share, want)
        if any(host):
            raise BsrError("inputs, uv and reg must all be CUDA tensors (device path) or all host arrays (host path)")
        import torch
        for name, a in (("inputs", inputs), ("uv", uv), ("reg", reg)):
            if a is None:
                continue
            if not getattr(a, "is_cuda", False):
                raise BsrError("%s: pass CUDA tensors (device path) or NumPy arrays (host path)" % name)
            if a.device.index != self.device:
                raise BsrError("%s lives on cuda:%s, this generator on cuda:%d" % (name, a.device.index, self.device))
        n = self._check_shapes(tuple(inputs.shape), tuple(uv.shape), None if reg is None else tuple(reg.shape), frame)
        inputs = inputs.contiguous().float()
        uv = uv.contiguous().float()
        dev = inputs.device
        out = {
            "gs": torch.empty((n, IMG, IMG, 1), device=dev) if "gs" in want else None,
            "con_rgb": torch.empty((n, IMG, IMG, 3), device=dev) if "con_rgb" in want else None,
            "mask22": torch.empty((n, IMG, IMG, 3), device=dev) if "mask22" in want else None,
            "dif": torch.empty((n, IMG, IMG, 1), device=dev) if "dif" in want else None,
        }
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        if self.variant == "gsc":
            rc = self._lib.bsr_forward_gsc(self._h, _ptr(inputs), _ptr(uv), n, _ptr(out["gs"]), _ptr(out["con_rgb"]),
                                           _ptr(out["mask22"]), _ptr(out["dif"]), stream)
        else:
            reg = reg.contiguous().float()
            rc = self._lib.bsr_forward_tsm(self._h, _ptr(inputs), _ptr(uv), _ptr(reg), n // frame, frame, int(share),
                                           _ptr(out["gs"]), _ptr(out["con_rgb"]), _ptr(out["mask22"]), _ptr(out["dif"]),
                                           stream)
        self._check(rc)
        return out["gs"], out["con_rgb"], out["mask22"], out["dif"]

    def _check_shapes(self, s_img, s_uv, s_reg, frame):
        if len(s_img) != 4 or s_img[1:] != (IMG, IMG, 3):
            raise ValueError("inputs must be [N,256,256,3] NHWC, got %r" % (s_img,))
        if s_uv != s_img:
            raise ValueError("uv must be [N,256,256,3], got %r" % (s_uv,))
        n = s_img[0]
        if n <= 0:
            raise ValueError("empty batch")
        if self.variant == "tsm":
            if frame is None or frame <= 0:
                raise ValueError("TSM variant needs frame > 0")
            if s_reg != (n, IMG, IMG, 6):
                raise ValueError("reg must be [N,256,256,6], got %r" % (s_reg,))
            if n % frame:
                raise ValueError("batch %d is not a multiple of frame %d (model_with_TSM.py:218)" % (n, frame))
        return n

    def _call_host(self, inputs, uv, reg, frame, share, want):
        n = self._check_shapes(inputs.shape, uv.shape, None if reg is None else reg.shape, frame)
        f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        inputs, uv = f32(inputs), f32(uv)
        out = {
            "gs": np.empty((n, IMG, IMG, 1), np.float32) if "gs" in want else None,
            "con_rgb": np.empty((n, IMG, IMG, 3), np.float32) if "con_rgb" in want else None,
            "mask22": np.empty((n, IMG, IMG, 3), np.float32) if "mask22" in want else None,
            "dif": np.empty((n, IMG, IMG, 1), np.float32) if "dif" in want else None,
        }
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        if self.variant == "gsc":
            rc = self._lib.bsr_forward_gsc_host(self._h, p(inputs), p(uv), n, p(out["gs"]), p(out["con_rgb"]),
                                                p(out["mask22"]), p(out["dif"]))
        else:
            reg = f32(reg)
            rc = self._lib.bsr_forward_tsm_host(self._h, p(inputs), p(uv), p(reg), n // frame, frame, int(share),
                                                p(out["gs"]), p(out["con_rgb"]), p(out["mask22"]), p(out["dif"]))
        self._check(rc)
        return out["gs"], out["con_rgb"], out["mask22"], out["dif"]

    def forward_host_ptrs(self, img_ptr, uv_ptr, reg_ptr, n, frame, share, gs_ptr, rgb_ptr, m22_ptr, dif_ptr):
        """Host path on raw (e.g. pinned) pointers; used by bench.py's e2e leg."""
        v = lambda x: None if not x else ctypes.c_void_p(x)
        if self.variant == "gsc":
            rc = self._lib.bsr_forward_gsc_host(self._h, v(img_ptr), v(uv_ptr), n, v(gs_ptr), v(rgb_ptr), v(m22_ptr), v(dif_ptr))
        else:
            rc = self._lib.bsr_forward_tsm_host(self._h, v(img_ptr), v(uv_ptr), v(reg_ptr), n // frame, frame, int(share),
                                                v(gs_ptr), v(rgb_ptr), v(m22_ptr), v(dif_ptr))
        self._check(rc)

    # -- chunk entry (SURVEY.md 8a row 0 + row 13) ---------------------------------------------
    def forward_chunk(self, chunk, frame: Optional[int] = None, share=True, want_raw: bool = False):
        """Body of the reference's test steps around the generator call (train_test_GSC.py:415-422, 802-809, 866-873,
        896-903; train_with_TSM.py:671-678, 713-720): split the dataset chunk [F,256,256,C] (C = 13 / 16 / 17), run the
        generator, ``mask_pred = dif * face``, ``rgb = clip(con_rgb, 0, 1)``.  ``chunk`` is a CUDA tensor (device path,
        returns CUDA tensors) or a NumPy array (uploaded once, returns NumPy).  Returns ``(rgb, mask_pred)`` or, with
        ``want_raw``, ``(rgb, mask_pred, gs, mask22)``."""
        import torch
        is_np = isinstance(chunk, np.ndarray)
        if not is_np and not getattr(chunk, "is_cuda", False) and (hasattr(chunk, "numpy") or hasattr(chunk, "__array__")):
            chunk, is_np = _as_host_array(chunk), True             # TF EagerTensor and friends
        t = torch.from_numpy(np.ascontiguousarray(chunk, dtype=np.float32)).cuda(self.device) if is_np else chunk
        if t.dim() != 4 or tuple(t.shape[1:3]) != (IMG, IMG) or int(t.shape[3]) not in CHUNK_LAYOUTS:
            raise ValueError("chunk must be [F,256,256,C] with C in {13, 16, 17}, got %r" % (tuple(t.shape),))
        if not t.is_cuda:
            raise BsrError("pass a CUDA tensor (device path) or a NumPy array")
        if t.device.index != self.device:
            raise BsrError("chunk lives on cuda:%s, this generator on cuda:%d" % (t.device.index, self.device))
        n = int(t.shape[0])
        if n <= 0:
            raise ValueError("empty chunk")
        if self.variant == "tsm":
            if frame is None or frame <= 0:
                raise ValueError("TSM variant needs frame > 0")
            if n % frame:
                raise ValueError("chunk of %d frames is not a multiple of frame %d (model_with_TSM.py:218)" % (n, frame))
        share = bool(share.item() if hasattr(share, "item") else share)
        t = t.contiguous().float()
        dev = t.device
        rgb = torch.empty((n, IMG, IMG, 3), device=dev)
        mp = torch.empty((n, IMG, IMG, 1), device=dev)
        gs = torch.empty((n, IMG, IMG, 1), device=dev) if want_raw else None
        m22 = torch.empty((n, IMG, IMG, 3), device=dev) if want_raw else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        self._check(self._lib.bsr_forward_chunk(self._h, _ptr(t), n, CHUNK_LAYOUTS[int(t.shape[3])][0], int(frame or 1),
                                                int(share), _ptr(rgb), _ptr(mp), _ptr(gs), _ptr(m22), stream))
        outs = (rgb, mp, gs, m22) if want_raw else (rgb, mp)
        if is_np:
            torch.cuda.synchronize(dev)
            return tuple(o.cpu().numpy() for o in outs)
        return outs

    # -- compact host I/O -----------------------------------------------
    def forward_compact(self, img_u8, uv32, reg32=None, frame: Optional[int] = None, share=True,
                        want=("rgb_u8", "dif_f16")):
        """Forward fed with the bytes the dataset really holds (dataset.py:119,159 read uint8 PNGs and
        divide by 255.; model.py:237 / warp.py:137 consume uv / reg only after the resize to 32x32).

        img_u8 [N,256,256,3] uint8; uv32 [N,32,32,3] fp32 (= ``downsample8(uv)``); reg32 [N,32,32,6] (TSM).
        ``want`` picks outputs from gs, con_rgb, mask22, dif (fp32 as in ``__call__``), rgb_u8
        (= rint(clip(con_rgb,0,1)*255)) and dif_f16.  Returns a dict of NumPy arrays.
        """
        img_u8 = np.ascontiguousarray(img_u8)
        if img_u8.dtype != np.uint8 or img_u8.ndim != 4 or img_u8.shape[1:] != (IMG, IMG, 3):
            raise ValueError("img_u8 must be uint8 [N,256,256,3], got %s %r" % (img_u8.dtype, img_u8.shape))
        n = img_u8.shape[0]
        if n <= 0:
            raise ValueError("empty batch")
        uv32 = np.ascontiguousarray(uv32, dtype=np.float32)
        if uv32.shape != (n, FEAT, FEAT, 3):
            raise ValueError("uv32 must be [N,32,32,3], got %r" % (uv32.shape,))
        if self.variant == "tsm":
            if frame is None or frame <= 0:
                raise ValueError("TSM variant needs frame > 0")
            if reg32 is None or tuple(np.shape(reg32)) != (n, FEAT, FEAT, 6):
                raise ValueError("reg32 must be [N,32,32,6]")
            if n % frame:
                raise ValueError("batch %d is not a multiple of frame %d (model_with_TSM.py:218)" % (n, frame))
            reg32 = np.ascontiguousarray(reg32, dtype=np.float32)
        unknown = set(want) - {"gs", "con_rgb", "mask22", "dif", "rgb_u8", "dif_f16"}
        if unknown:
            raise ValueError("unknown outputs %r" % sorted(unknown))
        spec = {"gs": ((n, IMG, IMG, 1), np.float32), "con_rgb": ((n, IMG, IMG, 3), np.float32),
                "mask22": ((n, IMG, IMG, 3), np.float32), "dif": ((n, IMG, IMG, 1), np.float32),
                "rgb_u8": ((n, IMG, IMG, 3), np.uint8), "dif_f16": ((n, IMG, IMG, 1), np.float16)}
        out = {k: (np.empty(*spec[k]) if k in want else None) for k in spec}
        p = lambda a: None if a is None else a.ctypes.data_as(ctypes.c_void_p)
        outs = (p(out["gs"]), p(out["con_rgb"]), p(out["mask22"]), p(out["dif"]), p(out["rgb_u8"]), p(out["dif_f16"]))
        self.forward_compact_ptrs(p(img_u8), p(uv32), p(reg32), n, frame, share, *outs)
        return {k: v for k, v in out.items() if v is not None}

    def forward_compact_ptrs(self, img_u8, uv32, reg32, n, frame, share, gs, rgb, m22, dif, rgb_u8, dif_f16):
        """Compact host path on raw (e.g. pinned) pointers (ints or c_void_p); used by bench.py."""
        v = lambda x: x if isinstance(x, ctypes.c_void_p) or x is None else (ctypes.c_void_p(x) if x else None)
        if self.variant == "gsc":
            rc = self._lib.bsr_forward_gsc_host_compact(self._h, v(img_u8), v(uv32), n, v(gs), v(rgb), v(m22), v(dif),
                                                        v(rgb_u8), v(dif_f16))
        else:
            share = bool(share.item() if hasattr(share, "item") else share)
            rc = self._lib.bsr_forward_tsm_host_compact(self._h, v(img_u8), v(uv32), v(reg32), n // frame, frame,
                                                        int(share), v(gs), v(rgb), v(m22), v(dif), v(rgb_u8), v(dif_f16))
        self._check(rc)

    # -- ShareLayer on its own (model_with_TSM.py:204-229) --------------------------------------
    def share_layer(self, x, reg, frame: int, share=True):
        """``ShareLayer.call(x, reg, frame, share)``: x [n,32,32,C] and reg [n,256,256,6] CUDA fp32 -> [n,32,32,2C]."""
        import torch
        self._same_device(x=x, reg=reg)
        n, c = int(x.shape[0]), int(x.shape[3])
        if tuple(x.shape[1:3]) != (FEAT, FEAT) or tuple(reg.shape) != (n, IMG, IMG, 6):
            raise ValueError("x must be [n,32,32,C] and reg [n,256,256,6]")
        if n % frame:
            raise ValueError("batch %d is not a multiple of frame %d (model_with_TSM.py:218)" % (n, frame))
        share = bool(share.item() if hasattr(share, "item") else share)
        out = torch.empty((n, FEAT, FEAT, 2 * c), dtype=torch.float32, device=x.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        self._check(self._lib.bsr_share_layer(self._h, _ptr(x.contiguous().float()), _ptr(reg.contiguous().float()), n, c,
                                              int(frame), int(share), _ptr(out), stream))
        return out

    # -- caller glue (train_test_GSC.py:808-809, 711-718) ------------------------------------
    def caller_glue(self, con_rgb, dif, face):
        import torch
        n = con_rgb.shape[0]
        self._same_device(con_rgb=con_rgb, dif=dif, face=face)
        # dense outputs: the kernel writes contiguous NHWC whatever the strides of the inputs were
        rgb_c = torch.empty(tuple(con_rgb.shape), dtype=torch.float32, device=con_rgb.device)
        mask_pred = torch.empty(tuple(dif.shape), dtype=torch.float32, device=con_rgb.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(con_rgb.device).cuda_stream)
        self._check(self._lib.bsr_caller_glue(self._h, _ptr(con_rgb.contiguous().float()), _ptr(dif.contiguous().float()),
                                              _ptr(face.contiguous().float()), n, _ptr(rgb_c), _ptr(mask_pred), stream))
        return rgb_c, mask_pred

    def _same_device(self, **tensors):
        for name, a in tensors.items():
            if not getattr(a, "is_cuda", False) or a.device.index != self.device:
                raise BsrError("%s must be a CUDA tensor on cuda:%d" % (name, self.device))

    def composite(self, pred, inp, m):
        import torch
        self._same_device(pred=pred, inp=inp, m=m)
        out = torch.empty(tuple(pred.shape), dtype=torch.float32, device=pred.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(pred.device).cuda_stream)
        self._check(self._lib.bsr_composite(self._h, _ptr(pred.contiguous().float()), _ptr(inp.contiguous().float()),
                                            _ptr(m.expand_as(pred).contiguous().float()), pred.numel(), _ptr(out), stream))
        return out

    # -- test_step post-processing (train_test_GSC.py:436-725) -------------------------------
    MASK_KINDS = ("face_hair", "face", "mouth", "nose", "eyebrow", "eye", "glasses")

    def postprocess_ucb(self, img, gt, con_rgb, dif, sizes, masks, want_metrics: bool = True):
        """Device post-processing of the UCB test step for n samples (frame 0 of n chunks): CUDA tensors img / gt /
        con_rgb [n,256,256,3], dif [n,256,256,1], sizes [n] int32, masks [n,7,256,256] uint8 in MASK_KINDS order.
        Returns (final [n,256,256,3], detected [n,256,256], metrics [n,2] = ssim, psnr or None)."""
        import torch
        self._same_device(img=img, gt=gt, con_rgb=con_rgb, dif=dif, sizes=sizes, masks=masks)
        n = int(img.shape[0])
        if tuple(img.shape) != (n, IMG, IMG, 3) or tuple(gt.shape) != (n, IMG, IMG, 3) or tuple(con_rgb.shape) != (n, IMG, IMG, 3):
            raise ValueError("img, gt, con_rgb must be [n,256,256,3]")
        if dif.numel() != n * IMG * IMG or tuple(masks.shape) != (n, 7, IMG, IMG) or masks.dtype != torch.uint8:
            raise ValueError("dif must be [n,256,256,1] and masks uint8 [n,7,256,256]")
        sizes = sizes.to(torch.int32).contiguous()
        if sizes.numel() != n or int(sizes.min()) < 1 or int(sizes.max()) > IMG:
            raise ValueError("sizes must be n values in [1, 256]")
        dev = img.device
        final = torch.empty((n, IMG, IMG, 3), dtype=torch.float32, device=dev)
        det = torch.empty((n, IMG, IMG), dtype=torch.float32, device=dev)
        met = torch.empty((n, 2), dtype=torch.float32, device=dev) if want_metrics else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        f = lambda t: t.contiguous().float()
        self._check(self._lib.bsr_postprocess_ucb(self._h, n, _ptr(f(img)), _ptr(f(gt)), _ptr(f(con_rgb)), _ptr(f(dif)),
                                                  _ptr(sizes), _ptr(masks.contiguous()), _ptr(final), _ptr(det), _ptr(met), stream))
        return final, det, met

    # -- introspection -----------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self._lib.bsr_launch_count(self._h))

    def check(self) -> None:
        """Synchronise the last forward and raise BsrError if a device watchdog fired (bsr_check)."""
        self._check(self._lib.bsr_check(self._h))

    def plan_counters(self) -> Dict[str, int]:
        """Launch-plan counters of the last forward (bsr_plan_counter)."""
        return {k: int(self._lib.bsr_plan_counter(self._h, i)) for i, k in enumerate(PLAN_COUNTERS)}

    def workspace_bytes(self) -> int:
        return int(self._lib.bsr_workspace_bytes(self._h))

    def debug_read(self, name: str) -> np.ndarray:
        n = ctypes.c_size_t()
        self._check(self._lib.bsr_debug_read(self._h, name.encode(), None, 0, ctypes.byref(n)))
        out = np.empty(n.value, np.float32)
        self._check(self._lib.bsr_debug_read(self._h, name.encode(), out.ctypes.data_as(ctypes.c_void_p), n.value,
                                             ctypes.byref(n)))
        return out

    def layer_times(self):
        names = (ctypes.c_char_p * 512)()
        ms = (ctypes.c_float * 512)()
        k = self._lib.bsr_layer_times(self._h, names, ms, 512)
        return [(names[i].decode(), float(ms[i])) for i in range(k)]

    def _check(self, rc):
        if rc != 0:
            raise BsrError("libbsr error %d: %s" % (rc, self._lib.bsr_last_error(self._h).decode()))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bsr_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
