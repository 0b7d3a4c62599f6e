"""Python owner of one libfdsr context: device memory and streams come from torch, compute from
the C ABI.  One Engine per (process, GPU)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import FdsrConfig, FdsrError, FdsrOverflowError

_DTYPES = {"auto": _lib.DTYPE_FP16, "fp16": _lib.DTYPE_FP16, "float16": _lib.DTYPE_FP16, "bf16": _lib.DTYPE_BF16, "bfloat16": _lib.DTYPE_BF16,
           "fp32": _lib.DTYPE_FP32, "float32": _lib.DTYPE_FP32}  # fp32 = CUDA-core parity mode (slow)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class Engine:
    def __init__(self, unet_cfg: dict, device, dtype: str = "fp16"):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise FdsrError("fastdiffsr_b200 runs on CUDA (sm_100a) devices only; there is no CPU fallback")
        mults = list(unet_cfg["channel_multiplier"])
        cfg = FdsrConfig()
        cfg.in_channel = unet_cfg["in_channel"]
        cfg.out_channel = unet_cfg["out_channel"]
        cfg.inner_channel = unet_cfg["inner_channel"]
        cfg.norm_groups = unet_cfg.get("norm_groups") or 32
        cfg.n_levels = len(mults)
        for i, m in enumerate(mults):
            cfg.channel_mults[i] = m
        cfg.res_blocks = unet_cfg["res_blocks"]
        cfg.dtype = _DTYPES[dtype]
        # which_model_G: "fastdiffsr" (default) or "ddpm" = the SR3 baseline, whose ResnetBlocks carry SelfAttention
        # where the resolution of the *configured* image_size is in attn_res (ddpm_modules/unet.py:184, 211)
        self.model = unet_cfg.get("model", "fastdiffsr")
        if self.model not in ("fastdiffsr", "ddpm"):
            raise FdsrError(f"unknown model {self.model!r}")
        cfg.model = _lib.MODEL_SR3 if self.model == "ddpm" else _lib.MODEL_FASTDIFFSR
        cfg.attn_levels = 0
        if self.model == "ddpm":
            res = int(unet_cfg.get("image_size", 256))
            for i in range(len(mults)):
                if res in list(unet_cfg["attn_res"]):
                    cfg.attn_levels |= 1 << i
                res //= 2
        self.dtype = dtype
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = self.lib.fdsr_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            raise FdsrError("fdsr_create: " + self.lib.fdsr_global_error().decode())
        self.T = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.fdsr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc < 0:
            cls = FdsrOverflowError if rc == _lib.E_OVERFLOW else FdsrError
            raise cls(f"{what}: {self.lib.fdsr_last_error(self._h).decode()} (code {rc})")
        return rc

    def check_overflow(self):
        """fp16 mode: synchronise the current stream and raise FdsrOverflowError if an activation left the fp16 range
        since the last check (the value was stored saturated; the result is not the network's output)."""
        if self.dtype not in ("fp16", "float16"):
            return
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_check_overflow(self._h, self._stream()), "fdsr_check_overflow")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # ---- weights / schedule
    def load_state_dict(self, sd: dict):
        """Tensors that already live on this engine's GPU are handed over as device pointers (fdsr_load_weights_dev);
        anything else goes through host memory (fdsr_load_weights)."""
        items = [(k, v.detach()) for k, v in sd.items() if k.startswith("denoise_fn.")]
        on_dev = bool(items) and all(v.device == self.device for _, v in items)
        items = [(k, v.to(torch.float32).contiguous() if on_dev else v.to("cpu", torch.float32).contiguous())
                 for k, v in items]
        n = len(items)
        names = (C.c_char_p * n)(*[k.encode() for k, _ in items])
        ptrs = (C.c_void_p * n)(*[v.data_ptr() for _, v in items])
        numels = (C.c_int64 * n)(*[v.numel() for _, v in items])
        with torch.cuda.device(self.device):
            if on_dev:
                self._check(self.lib.fdsr_load_weights_dev(self._h, names, ptrs, numels, n, self._stream()),
                            "fdsr_load_weights_dev")
            else:
                self._check(self.lib.fdsr_load_weights(self._h, names, ptrs, numels, n), "fdsr_load_weights")

    def set_schedule(self, betas):
        b = np.ascontiguousarray(np.asarray(betas, dtype=np.float64))
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_set_schedule(self._h, b.ctypes.data_as(C.POINTER(C.c_double)), len(b)),
                        "fdsr_set_schedule")
        self.T = len(b)

    def table(self, name: str):
        buf = (C.c_double * (self.T + 1))()
        n = self._check(self.lib.fdsr_get_table(self._h, name.encode(), buf, self.T + 1), "fdsr_get_table")
        return np.array(buf[:n], dtype=np.float64)

    # ---- compute
    def _img(self, t, name):
        if t.device != self.device or t.dtype != torch.float32 or t.dim() != 4 or t.shape[1] != 3:
            raise FdsrError(f"{name} must be a (B,3,H,W) fp32 tensor on {self.device}")
        return t.contiguous()

    def unet_forward(self, cond, x_t, t: int):
        cond, x_t = self._img(cond, "cond"), self._img(x_t, "x_t")
        B, _, H, W = cond.shape
        out = torch.empty_like(cond)
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_unet_forward(self._h, _ptr(cond), _ptr(x_t), t, _ptr(out), B, H, W,
                                                   self._stream()), "fdsr_unet_forward")
        return out

    def posterior_step(self, x_t, eps, z, t: int):
        x_t, eps = x_t.contiguous(), eps.contiguous()
        z = z.contiguous() if z is not None else None
        out = torch.empty_like(x_t)
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_posterior_step(self._h, _ptr(x_t), _ptr(eps), _ptr(z), t, _ptr(out),
                                                     x_t.numel(), self._stream()), "fdsr_posterior_step")
        return out

    def trace_frames(self):
        return self.lib.fdsr_trace_frames(self._h)

    def sample(self, cond, noise=None, seed: int = 0, trace: bool = False, image_offset: int = 0):
        """T-step sampling of a batch.  `image_offset` = global index of cond[0] in the whole job: the built-in noise of
        an image depends on (seed, global index, step) only, so shards of a job reproduce the unsharded result."""
        self._check(self.lib.fdsr_set_image_offset(self._h, int(image_offset)), "fdsr_set_image_offset")
        cond = self._img(cond, "cond")
        B, _, H, W = cond.shape
        if noise is not None:
            if tuple(noise.shape) != (self.T, B, 3, H, W) or noise.dtype != torch.float32 or noise.device != self.device:
                raise FdsrError(f"noise must be ({self.T},{B},3,{H},{W}) fp32 on {self.device}")
            noise = noise.contiguous()
        out = torch.empty_like(cond)
        tr = torch.empty((B, self.trace_frames(), 3, H, W), device=self.device, dtype=torch.float32) if trace else None
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_sample(self._h, _ptr(cond), _ptr(noise), seed, _ptr(out), _ptr(tr), B, H, W,
                                             self._stream()), "fdsr_sample")
        return (out, tr) if trace else out

    def bicubic_u8(self, lr_u8, H: int, W: int, want_u8=True, want_cond=True):
        """lr_u8: (B,h,w,3) uint8 on device -> (u8 (B,H,W,3) | None, cond (B,3,H,W) fp32 | None)."""
        if lr_u8.dtype != torch.uint8 or lr_u8.dim() != 4 or lr_u8.shape[3] != 3 or lr_u8.device != self.device:
            raise FdsrError("lr must be (B,h,w,3) uint8 on the engine's device")
        lr_u8 = lr_u8.contiguous()
        B, h, w, _ = lr_u8.shape
        o8 = torch.empty((B, H, W, 3), dtype=torch.uint8, device=self.device) if want_u8 else None
        oc = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device) if want_cond else None
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_bicubic_u8(self._h, _ptr(lr_u8), B, h, w, H, W, _ptr(o8), _ptr(oc),
                                                 self._stream()), "fdsr_bicubic_u8")
        return o8, oc

    def debug_noise(self, B: int, H: int, W: int, seed: int, stream_id: int, image_offset: int = 0):
        """(B,3,H,W) N(0,1) values of the built-in generator for step stream `stream_id` (T: x_T, t: z of step t)."""
        out = torch.empty((B, 3, H, W), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_debug_noise(self._h, _ptr(out), B, H, W, int(seed), int(image_offset),
                                                  int(stream_id), self._stream()), "fdsr_debug_noise")
        return out

    def graph_captures(self):
        return int(self.lib.fdsr_graph_captures(self._h))

    def super_resolve_u8_host(self, lr_host: np.ndarray, H: int, W: int, noise=None, seed: int = 0, out=None,
                              image_offset: int = 0):
        """Host uint8 (B,h,w,3) -> host fp32 (B,3,H,W): H2D, bicubic, T-step sampling, D2H.
        `out`: optional caller-owned C-contiguous float32 (B,3,H,W) array to fill (avoids a fresh allocation per call)."""
        lr_host = np.ascontiguousarray(lr_host, dtype=np.uint8)
        B, h, w, _ = lr_host.shape
        if out is None:
            out = np.empty((B, 3, H, W), dtype=np.float32)
        elif out.shape != (B, 3, H, W) or out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]:
            raise FdsrError(f"out must be a C-contiguous float32 array of shape {(B, 3, H, W)}")
        self._check(self.lib.fdsr_set_image_offset(self._h, int(image_offset)), "fdsr_set_image_offset")
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_super_resolve_u8(self._h, lr_host.ctypes.data_as(C.c_void_p), B, h, w, H, W,
                                                       _ptr(noise), seed, out.ctypes.data_as(C.c_void_p),
                                                       self._stream()), "fdsr_super_resolve_u8")
        return out

    def super_resolve_u8_submit(self, slot: int, lr_host: np.ndarray, H: int, W: int, noise=None, seed: int = 0,
                                image_offset: int = 0):
        """Pipelined host path: enqueue one batch into `slot` (0 / 1) and return; see super_resolve_u8_wait."""
        lr_host = np.ascontiguousarray(lr_host, dtype=np.uint8)
        B, h, w, _ = lr_host.shape
        self._check(self.lib.fdsr_set_image_offset(self._h, int(image_offset)), "fdsr_set_image_offset")
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_super_resolve_u8_submit(self._h, slot, lr_host.ctypes.data_as(C.c_void_p), B, h, w, H, W,
                                                              _ptr(noise), seed, self._stream()),
                        "fdsr_super_resolve_u8_submit")
        return (B, 3, H, W)

    def super_resolve_u8_wait(self, slot: int, out: np.ndarray):
        """Block until the batch in `slot` is on the host; fills the C-contiguous float32 array `out`."""
        if out.dtype != np.float32 or not out.flags["C_CONTIGUOUS"]:
            raise FdsrError("out must be a C-contiguous float32 array")
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_super_resolve_u8_wait(self._h, slot, out.ctypes.data_as(C.c_void_p)),
                        "fdsr_super_resolve_u8_wait")
        return out

    def sse_u8(self, a, b):
        a, b = self._img(a, "a"), self._img(b, "b")
        B, _, H, W = a.shape
        out = torch.empty(B, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_sse_u8(self._h, _ptr(a), _ptr(b), B, H, W, _ptr(out), self._stream()),
                        "fdsr_sse_u8")
        return out

    def metrics_u8(self, a, b, scale: float = 4.0):
        """Per-image (mse, psnr, ssim, ergas) of `a` (SR or bicubic image) against `b` (HR), both (B,3,H,W) fp32
        in [-1,1]: the reference's evaluation block (sr_mfe.py:315-345) on the device.  Returns (B,4) float64."""
        a, b = self._img(a, "a"), self._img(b, "b")
        B, _, H, W = a.shape
        out = torch.empty((B, 4), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_metrics_u8(self._h, _ptr(a), _ptr(b), B, H, W, float(scale), _ptr(out),
                                                 self._stream()), "fdsr_metrics_u8")
        return out

    # ---- hooks
    def tensor_names(self):
        n = self.lib.fdsr_debug_num_tensors(self._h)
        return [self.lib.fdsr_debug_tensor_name(self._h, i).decode() for i in range(n)]

    def read_tensor(self, name: str, B: int, max_elems: int):
        buf = torch.empty(max_elems, dtype=torch.float32, device=self.device)
        c, h, w = C.c_int32(), C.c_int32(), C.c_int32()
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_debug_read_tensor(self._h, name.encode(), _ptr(buf), max_elems, C.byref(c),
                                                        C.byref(h), C.byref(w), self._stream()), "fdsr_debug_read_tensor")
        n = B * c.value * h.value * w.value
        return buf[:n].view(B, c.value, h.value, w.value)

    def op_flops_executed(self):
        """Executed conv FLOPs per op (phase-decomposed upsample convs at 4/9 of their nine-tap cost)."""
        n = self.lib.fdsr_debug_num_ops(self._h)
        return [float(self.lib.fdsr_debug_op_flops_executed(self._h, i)) for i in range(n)]

    def profile_unet(self, t: int, reps: int = 3):
        """[(op name, mean ms, algorithmic conv FLOPs)] for one UNet evaluation at the current shape."""
        n = self.lib.fdsr_debug_num_ops(self._h)
        ms = (C.c_float * n)()
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_debug_profile_unet(self._h, t, reps, ms, n, self._stream()),
                        "fdsr_debug_profile_unet")
        return [(self.lib.fdsr_debug_op_name(self._h, i).decode(), float(ms[i]),
                 float(self.lib.fdsr_debug_op_flops(self._h, i))) for i in range(n)]

    def role_cycles(self, op: int, t: int = 0):
        """(n_cta, 5 roles, 8 slots) cycle counters of conv op `op` (FDSR_PROFILE builds only)."""
        n = 148 * 40 * 2
        buf = (C.c_int64 * n)()
        with torch.cuda.device(self.device):
            m = self._check(self.lib.fdsr_debug_role_cycles(self._h, op, t, buf, n, self._stream()),
                            "fdsr_debug_role_cycles")
        return np.array(buf[:m], dtype=np.int64).reshape(-1, 5, 8)

    def timeline(self, t: int = 0, reps: int = 3):
        """(n_ops, n_sms, 5 roles, 8 slots): the same counters of every conv op inside a running UNet evaluation
        (FDSR_PROFILE builds only; tools/timeline.py)."""
        nops = int(self.lib.fdsr_debug_num_ops(self._h))
        nsm = torch.cuda.get_device_properties(self.device).multi_processor_count
        out = np.zeros((nops, nsm, 5, 8), dtype=np.int64)
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_debug_timeline(self._h, t, reps, out.ctypes.data_as(C.POINTER(C.c_int64)), out.size,
                                                     self._stream()), "fdsr_debug_timeline")
        return out

    def launch_count(self):
        return int(self.lib.fdsr_launch_count(self._h))

    def unet_flops(self):
        return float(self.lib.fdsr_unet_flops(self._h))

    def reserve(self, B, H, W):
        with torch.cuda.device(self.device):
            self._check(self.lib.fdsr_reserve(self._h, B, H, W), "fdsr_reserve")

    def workspace_bytes(self):
        return int(self.lib.fdsr_workspace_bytes(self._h))

    def set_use_graph(self, enable: bool):
        self.lib.fdsr_set_use_graph(self._h, 1 if enable else 0)
