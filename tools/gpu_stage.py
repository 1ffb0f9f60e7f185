#!/usr/bin/env python
"""Staged bring-up on the GPU box.  Each stage runs in its own process under a timeout so that a faulty
kernel cannot take the rest of the session with it:   python tools/gpu_stage.py all   (or one stage name).
Writes one log per stage to gpurun_out/stage_<name>.log."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "gpurun_out")
STAGES = ["selftest", "weights", "fp32", "probe", "tc", "big"]


def tf32_rna(x):
    import numpy as np
    b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)
    return b.view(np.float32)


def stage_selftest():
    import torch
    from simhand_b200 import _lib
    lib = _lib.load()
    for which, name in enumerate(["sqrt", "sqrt2", "div21", "divw"]):
        out = torch.zeros(8, dtype=torch.int64, device="cuda")
        t0 = time.time()
        _lib.check(lib.smh_selftest(which, out.data_ptr(), 8, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        v = out.cpu().tolist()
        print(f"{name}: tested {v[0]} bad {v[1]} first {v[2]:#x} max_ulp {v[3]}  ({time.time() - t0:.2f}s)")


def _golden(name):
    import numpy as np
    return dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))


def _dev_inputs(g):
    import torch
    dev = torch.device("cuda")
    j1, j2 = torch.from_numpy(g["joints1"]).to(dev), torch.from_numpy(g["joints2"]).to(dev)
    return torch.from_numpy(g["z1"]).to(dev), torch.from_numpy(g["z2"]).to(dev), j1[:, :, :2], j2[:, :, :2]


def stage_weights():
    import numpy as np
    from oracle import restate as R
    from simhand_b200 import ops
    for name in ("n3_uniform", "n64_hand", "n96_uniform", "n200_peclr", "n256_uniform"):
        g = _golden(name)
        z1, z2, a, b = _dev_inputs(g)
        pw, nw = ops.mpjpe_weights(a, b)
        up = R.ulp_distance(pw.cpu().numpy(), g["pos_w"])
        un = R.ulp_distance(nw.cpu().numpy(), g["neg_w"])
        print(f"{name}: pos ulp max {up.max()} neg ulp max {un.max()} (nonzero {int((un > 0).sum())} of {un.size})")
        if un.max() > 0:
            idx = np.argwhere(un > 0)[:5]
            for i, j in idx:
                print("   ", i, j, nw[i, j].item(), g["neg_w"][i, j])


def _report_step(tag, loss, dz1, dz2, aux, g):
    import numpy as np
    from oracle import restate as R
    ref = float(g["loss_f64"])
    st = aux["stats"].cpu().numpy()
    c1, m1 = R.grad_metrics(dz1.cpu().numpy(), g["dz1_f64"])
    c2, m2 = R.grad_metrics(dz2.cpu().numpy(), g["dz2_f64"])
    print(f"{tag}: loss {float(loss):.8f} ref {ref:.8f} rel {abs(float(loss) - ref) / abs(ref):.2e} | "
          f"dz1 cos {c1:.8f} maxerr {m1:.2e} | dz2 cos {c2:.8f} maxerr {m2:.2e} | flags {st[3]} fail_site {st[6]}")


def stage_engine(engine):
    from simhand_b200 import ops
    for name in ("n64_hand", "n3_uniform", "n96_uniform", "n200_peclr", "n256_hand", "n130_hand_d64"):
        g = _golden(name)
        z1, z2, a, b = _dev_inputs(g)
        loss, dz1, dz2, aux = ops.run_step(z1, z2, a, b, 0.5, engine, True, return_aux=True)
        _report_step(f"{engine} {name}", loss, dz1, dz2, aux, g)


def bf16_rne(x):
    import numpy as np
    b = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    return b.view(np.float32)


def stage_probe():
    import numpy as np
    import torch
    from simhand_b200 import _lib, ops, synth
    lib = _lib.load()
    dev = torch.device("cuda")
    n = 128
    z1, z2, j1, j2 = synth.make_batch(n, 128, 3, "hand")
    ctx = ops.get_context(n, 128, 1, 0, dev)
    inp, keep = ops.make_inputs(z1.to(dev), z2.to(dev), j1.to(dev)[:, :, :2], j2.to(dev)[:, :, :2])
    ws = torch.empty(int(ctx.layout.ws_bytes), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.smh_prep(ctypes.byref(ctx.dims), ctypes.byref(inp), ws.data_ptr(), 0, st))
    zt_ptr = ws.data_ptr() + ctx.layout.off_zt
    zb_ptr = ws.data_ptr() + ctx.layout.off_zb
    zt32 = tf32_rna(torch.cat([z1, z2]).numpy())
    zfull = zt32.astype(np.float64)
    zb16 = bf16_rne(zt32).astype(np.float64)
    blk_a, blk_b = 0, 3
    A = zfull[blk_a * 64: blk_a * 64 + 128]
    B = zfull[blk_b * 64: blk_b * 64 + 64]
    Bb = zb16[blk_b * 64: blk_b * 64 + 64]
    s_ref = A @ B.T

    def run(params, tag=None):
        arr = (ctypes.c_uint32 * 16)(*params)
        s_out = torch.full((128, 64), float("nan"), device=dev)
        g_out = torch.zeros((128, 32), dtype=torch.int32, device=dev)
        dz_out = torch.full((128, 128), float("nan"), device=dev)
        fail = torch.zeros(1, dtype=torch.int32, device=dev)
        _lib.check(lib.smh_tc_probe(zt_ptr, zb_ptr, blk_a, blk_b, arr, s_out.data_ptr(), g_out.data_ptr(),
                                    dz_out.data_ptr(), fail.data_ptr(), st))
        torch.cuda.synchronize()
        s = s_out.cpu().numpy().astype(np.float64)
        g = g_out.cpu().numpy().view(np.uint32)
        glo = ((g & 0xFFFF) << 16).astype(np.uint32).view(np.float32)
        ghi = (g & 0xFFFF0000).astype(np.uint32).view(np.float32)
        gq = bf16_rne(s.astype(np.float32)).astype(np.float64)
        gread = np.empty((128, 64))
        gread[:, 0::2], gread[:, 1::2] = glo, ghi
        dz = dz_out.cpu().numpy().astype(np.float64)
        dz_ref = gq @ Bb
        e1 = np.abs(s - s_ref).max()
        eg = np.abs(gread - gq).max()
        e2 = np.abs(dz - dz_ref).max()
        if tag:
            np.savez(os.path.join(OUT, f"probe_{tag}.npz"), s=s, g=gread, dz=dz, s_ref=s_ref, dz_ref=dz_ref, B=Bb)
            print(f"[{tag}] dz min {np.nanmin(dz):.4f} max {np.nanmax(dz):.4f} nan {int(np.isnan(dz).sum())} | ref min "
                  f"{dz_ref.min():.4f} max {dz_ref.max():.4f} | dz[0,:4] {dz[0,:4]} ref {dz_ref[0,:4]}")
        return e1, eg, e2, int(fail.item())

    d = (ctypes.c_uint32 * 16)()
    lib.smh_tc_default_params(d)
    base = list(d)
    print("default params", base)
    e1, eg, e2, f = run(base, "default")
    print(f"default: max|S - ref| {e1:.3e}  max|G readback - bf16(S)| {eg:.3e}  max|dZ - ref| {e2:.3e}  fail {f}")
    if not (e2 < 1e-4):
        for lbo, sbo in ((8192, 1024), (1024, 8192), (8192, 128), (128, 8192), (16, 1024), (8192, 16), (4096, 1024)):
            for kstep in (2048, 1024, 256, 32):
                for colstep in (8, 16):
                    p = list(base)
                    p[9], p[10], p[11], p[12] = lbo, sbo, kstep, colstep
                    _, _, e2, f = run(p)
                    print(f"  MN-major bf16 lbo {lbo} sbo {sbo} kstep {kstep} colstep {colstep}: err {e2:.3e} fail {f}")


def stage_big():
    import torch
    from simhand_b200 import ops, synth
    dev = torch.device("cuda")
    z1, z2, j1, j2 = synth.make_batch(8192, 128, 5, "hand")
    a, b, c, d = z1.to(dev), z2.to(dev), j1.to(dev), j2.to(dev)
    for engine in ("tf32", "bf16", "fp32"):
        for it in range(3):
            torch.cuda.synchronize()
            t0 = time.time()
            loss, dz1, dz2, aux = ops.run_step(a, b, c[:, :, :2], d[:, :, :2], 0.5, engine, True, return_aux=True)
            torch.cuda.synchronize()
            st = aux["stats"].cpu().numpy()
            print(f"{engine} it{it}: loss {float(loss):.7f} |dz|max {float(dz1.abs().max()):.3e} "
                  f"fail_site {st[6]} wall {1e3 * (time.time() - t0):.2f} ms")


def main():
    os.makedirs(OUT, exist_ok=True)
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "all":
        rc = 0
        for s in STAGES:
            log = os.path.join(OUT, f"stage_{s}.log")
            t0 = time.time()
            with open(log, "w") as fh:
                p = subprocess.run(["timeout", "-k", "10", "240", sys.executable, __file__, s], stdout=fh,
                                   stderr=subprocess.STDOUT)
            print(f"== stage {s}: exit {p.returncode} in {time.time() - t0:.1f}s")
            sys.stdout.write(open(log).read()[-6000:])
            rc |= p.returncode
        sys.exit(1 if rc else 0)
    fn = {"selftest": stage_selftest, "weights": stage_weights, "fp32": lambda: stage_engine("fp32"),
          "probe": stage_probe, "tc": lambda: stage_engine("tf32"), "bf16": lambda: stage_engine("bf16"), "fp16": lambda: stage_engine("fp16"),
          "big": stage_big}[what]
    fn()


if __name__ == "__main__":
    main()
