"""CUDA ray caster for bench.py's whole-sequence workloads (synth_gpu.cu): input generator only.

    gs = GpuSynth(scene, device)                      # uploads the scene's primitives once
    flat, offs = gs.scan_batch(poses, seeds, sensor)  # torch float32 [total, 4] on the device + int64 offsets [n + 1]

Clouds come out in the CPU generator's order (azimuth-major) but are not bit-identical to it (libm vs CUDA
transcendentals): both arms of a benchmark must read the clouds made here.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import Sensor, _lib as _cpu_lib

_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        p = Path(__file__).resolve().parent / "libmlo_synth_cuda.so"
        if not p.exists():
            from .._build import build_synth_cuda
            build_synth_cuda()
        L = C.CDLL(str(p))
        L.synth_gpu_scan_batch.restype = C.c_int
        L.synth_gpu_scan_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


class GpuSynth:
    CAP = 8192   # primitives within range of one sensor pose (the street grid holds ~1000 within 120 m)

    def __init__(self, scene, device):
        import torch
        self.torch = torch
        self.dev = torch.device(device)
        L = _cpu_lib()
        L.synth_scene_export.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        nb, nc, ns = C.c_int(), C.c_int(), C.c_int()
        L.synth_scene_export(scene._h, C.byref(nb), None, C.byref(nc), None, C.byref(ns), None)
        b, c, s = np.empty((nb.value, 6), np.float32), np.empty((nc.value, 5), np.float32), np.empty((ns.value, 4), np.float32)
        L.synth_scene_export(scene._h, C.byref(nb), b.ctypes.data, C.byref(nc), c.ctypes.data, C.byref(ns), s.ctypes.data)
        self.boxes, self.cyls, self.sphs = (torch.from_numpy(x).to(self.dev) for x in (b, c, s))

    def scan_batch(self, poses, seeds, sensor: Sensor):
        """poses [n, 3, 4] sensor->world, seeds [n] -> (float32 [total, 4] x y z intensity, int64 offsets [n + 1])."""
        torch = self.torch
        n = len(poses)
        n_rays = sensor.n_beams * sensor.n_az
        dp = torch.from_numpy(np.ascontiguousarray(np.asarray(poses, dtype=np.float64)[:, :3, :4])).to(self.dev)
        ds = torch.from_numpy(np.ascontiguousarray(np.asarray(seeds, dtype=np.uint64).view(np.int64))).to(self.dev)
        lists = torch.empty((n, self.CAP), dtype=torch.int32, device=self.dev)
        counts = torch.empty((n,), dtype=torch.int32, device=self.dev)
        out = torch.empty((n, n_rays, 4), dtype=torch.float32, device=self.dev)
        st = torch.cuda.current_stream(self.dev).cuda_stream
        rc = _lib().synth_gpu_scan_batch(self.boxes.data_ptr(), len(self.boxes), self.cyls.data_ptr(), len(self.cyls),
                                         self.sphs.data_ptr(), len(self.sphs), dp.data_ptr(), ds.data_ptr(), n, sensor.n_beams,
                                         sensor.n_az, sensor.el_top_deg, sensor.el_bot_deg, sensor.max_range, sensor.noise_sigma,
                                         self.CAP, lists.data_ptr(), counts.data_ptr(), out.data_ptr(), C.c_void_p(st))
        if rc != 0:
            raise RuntimeError(f"synth_gpu_scan_batch: CUDA error {rc}")
        if int(counts.max()) > self.CAP:
            raise RuntimeError("GpuSynth: more primitives in range than CAP")
        mask = out[:, :, 3] >= 0
        offs = torch.zeros(n + 1, dtype=torch.int64, device=self.dev)
        offs[1:] = torch.cumsum(mask.sum(1), 0)
        return out[mask], offs
