#!/usr/bin/env python
"""bench.py — MPixels/s decoded on BASELINE.json's configs[1]:
   batch=1024 1920x1080 4:2:0 baseline JPEG, RST interval = 4 MCUs, per B200 (weak scaling).

  python bench.py --gpus N --steps K --warmup W          our arm (one rank per GPU, torchrun for N>1)
  python bench.py --impl reference --steps K --warmup W  the reference's CPU DecodeScanImg on the host cores

A "step" = one pass of the hot path over the whole batch: marker scan -> unstuff -> Huffman ->
dequant/IDCT/upsample/colour -> maps/statistics, i.e. everything CimgDecode::DecodeScanImg does
(ImgDecode.cpp:2723-3745), for every image of the batch.
  value : SOF pixels (X*Y, not padded) of all ranks / device time, inputs resident in HBM.
  e2e   : same metric through the one-call C-ABI jsgpu_decode_batch_host: pinned host bitstream in,
          every reference output (Y/Cb/Cr int16 maps, BGRA DIB, block-DC maps, MCU map, histo, stats)
          back in pinned host memory, H2D and D2H inside the timed region.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {
    # name: (batch, width, height, subsampling, quality, restart_interval)
    "cfg2": (1024, 1920, 1080, "420", 85, 4),
    "cfg1": (1, 640, 480, "444", 85, 80),
    "cfg3shard": (512, 3840, 2160, "420", 85, 8),     # one GPU's shard of config 3
    "cfg5": (512, 3840, 2160, "420", 85, 0),
}
METRIC = "MPixels/s decoded (bit-exact vs ref)"
UNIT = "MPix/s"


def make_batch(cfg, rank, nimg=None):
    from jpegsnoop_b200 import synth
    batch, w, h, ss, q, ri = CONFIGS[cfg]
    if nimg is not None:
        batch = nimg
    cfgnum = {"cfg1": 1, "cfg2": 2, "cfg3shard": 3, "cfg5": 5}[cfg]
    specs = [dict(width=w, height=h, subsampling=ss, quality=q, restart_interval=ri, optimize=False,
                  seed=1234 + cfgnum * 1000 + rank * batch + i) for i in range(batch)]
    buf, offs = synth.encode_batch(specs)
    return [buf[int(offs[i]):int(offs[i + 1])] for i in range(batch)], (w, h, ss, q, ri)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []; self.p = None; self.index = index

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # median of the upper half = clocks under load (idle samples before/after drag the plain median down)
        load = sm[len(sm) // 2:] if sm else []
        return {"sm_mhz": (load[len(load) // 2] if load else None), "sm_max_mhz": (max(mx) if mx else None),
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(jpegs, npix_per_img, budget_s=15.0, threads=None, extras=True):
    """The reference's own CPU DecodeScanImg (oracle/_ref, -DIDCT_FIXEDPT build) on the host cores,
    on a bounded sample of the same batch.  Falls back to the C port when _ref is absent."""
    from oracle_util import Oracle, ref_available
    kind = "reference" if ref_available("fixed") else "port"
    orc = Oracle("ref_fixed") if kind == "reference" else Oracle("port", idct_fixed=True)
    cores = threads or (os.cpu_count() or 1)
    # calibrate with one image per thread, then size the sample to the budget
    probe = [bytes(j) for j in jpegs[:min(len(jpegs), cores)]]
    t, errs = orc.bench(probe, threads=cores, reps=1)
    per_img = max(t / max(len(probe), 1) * min(cores, len(probe)), 1e-4)    # thread-seconds per image
    n = int(max(cores, min(len(jpegs), budget_s * cores / per_img)))
    sample = [bytes(j) for j in jpegs[:n]]
    t, errs = orc.bench(sample, threads=cores, reps=1)
    out = {"value": round(len(sample) * npix_per_img / t / 1e6, 2), "unit": UNIT, "cores": cores, "kind": kind,
           "sample": f"first {len(sample)} images of the batch, {cores} threads, one CimgDecode per thread, "
                     f"integer-IDCT build, {t:.2f} s wall", "err_lines": errs}
    if extras:      # also one thread, and the float-IDCT build (the reference's shipping default)
        try:
            k1 = [bytes(j) for j in jpegs[:max(1, min(len(jpegs), int(2.0 / per_img) or 1))]]
            t1, _ = orc.bench(k1, threads=1, reps=1)
            out["one_thread_value"] = round(len(k1) * npix_per_img / t1 / 1e6, 2)
            if kind == "reference" and ref_available("float"):
                of = Oracle("ref_float")
                kf = sample[:max(cores, len(sample) // 3)]
                tf, _ = of.bench(kf, threads=cores, reps=1)
                out["float_idct_build_value"] = round(len(kf) * npix_per_img / tf / 1e6, 2)
        except Exception as e:                      # the headline CPU figure above does not depend on these
            out["extras_error"] = str(e)[:200]
    return out, t, len(sample)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import __graft_entry__ as g
    g.build()
    cores = os.cpu_count() or 1
    nimg = max(cores, 64)
    jpegs, (w, h, ss, q, ri) = make_batch(args.config, 0, nimg=nimg)
    npix = w * h
    from oracle_util import Oracle, ref_available
    kind = "reference" if ref_available("fixed") else "port"
    orc = Oracle("ref_fixed") if kind == "reference" else Oracle("port", idct_fixed=True)
    sample = [bytes(j) for j in jpegs]
    for _ in range(args.warmup):
        orc.bench(sample[:cores], threads=cores, reps=1)
    t0 = time.time(); tot = 0.0
    for _ in range(args.steps):
        t, errs = orc.bench(sample, threads=cores, reps=1); tot += t
    val = args.steps * len(sample) * npix / tot / 1e6
    line = {"impl": "reference", "metric": METRIC, "value": round(val, 2), "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(tot / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "int32 IDCT + f32 colour", "data": "synthetic",
            "config": {"workload": f"batch={CONFIGS[args.config][0]} {w}x{h} 4:2:0 baseline, RST interval={ri} MCUs (BASELINE configs[1]); "
                                   f"each step = a bounded sample of {len(sample)} images of that batch"},
            "cpu_baseline": {"value": round(val, 2), "unit": UNIT, "cores": cores, "kind": kind,
                             "sample": f"{len(sample)} images per step, {cores} host threads, one CimgDecode per thread, integer-IDCT build"},
            "e2e": {"value": round(val, 2), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=list(CONFIGS))
    ap.add_argument("--batch", type=int, default=None, help="override images per GPU (debug only; invalidates the headline)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-histo", action="store_true", help="debug: skip m_anDhtHisto (invalidates the headline)")
    ap.add_argument("--no-mcu-map", action="store_true", help="debug: skip m_pMcuFileMap (invalidates the headline)")
    ap.add_argument("--huff-kernel", type=int, default=0)
    ap.add_argument("--idct-kernel", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        import __graft_entry__ as g
        g.build()
    if world > 1:
        dist.barrier()
    from jpegsnoop_b200 import BatchDecoder, _lib as B
    import ctypes as C

    # ---- synthetic batch of this rank (image-sharded: each rank owns its own images) ----------------
    t0 = time.time()
    jpegs, (w, h, ss, q, ri) = make_batch(args.config, rank, nimg=args.batch)
    nimg = len(jpegs)
    t_gen = time.time() - t0
    bd = BatchDecoder(device=local, huff_kernel=args.huff_kernel, idct_kernel=args.idct_kernel, want_histo=not args.no_histo, want_mcu_map=not args.no_mcu_map)
    tarr, darr, bits = BatchDecoder.prepare(jpegs)
    # ---- shared Huffman/quant tables: broadcast rank 0's table blob over NCCL (NVLink) ---------------
    if world > 1:
        from jpegsnoop_b200.shard import broadcast_tables
        tarr = broadcast_tables(tarr, src=0, device=torch.device("cuda", local))
    bd.set_tables(tarr)
    bd.plan(darr, bits.size)
    bd.upload(bits); bd.sync()

    # ---- device-resident timing ------------------------------------------------------------------------
    for _ in range(args.warmup):
        bd.decode()
    bd.sync()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start(); time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    stage_acc = np.zeros(5)
    bd.timer_start()
    for _ in range(args.steps):
        bd.decode()
    ms_total = bd.timer_stop()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    # per-stage times of the LAST step (CUDA events recorded inside jsgpu_batch_decode on the same stream)
    stage_ms = bd.stage_ms()
    launches = bd.launches() * args.steps
    status = sorted(set(int(l.status) for l in bd.refresh_layout()))
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    npx = torch.tensor([float(bd.nsof_pixels)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX); dist.all_reduce(npx, op=dist.ReduceOp.SUM)
    ms_max = float(t.item()); total_px = float(npx.item())
    value = total_px * args.steps / (ms_max / 1e3) / 1e6

    # ---- bit-exactness spot check against the CPU oracle on images spread over the batch (rank 0) -----------
    checked = []; parity_ok = True; orc = None
    if rank == 0:
        from oracle_util import Oracle, ref_available
        import jpeg_cases as JC
        orc = Oracle("ref_fixed") if ref_available("fixed") else Oracle("port", idct_fixed=True)
        WHAT = ("geom", "pix_y", "pix_cb", "pix_cr", "dib", "mcu_map", "blk_dc", "dht_histo")
        for i in sorted(set([0, nimg // 3, (2 * nimg) // 3, nimg - 1])):
            bad = JC.compare(orc.decode(bytes(jpegs[i])), bd.fetch(i), what=WHAT)
            checked.append(i); parity_ok = parity_ok and not bad

    # ---- end-to-end: host buffers in/out through the one-call C-ABI --------------------------------------
    e2e = None
    if not args.no_e2e:
        L = B.load()
        # host buffers should live on the NUMA node the GPU hangs off: run the e2e leg on that node's cores
        # (pinned pages are first-touch local) and give the process its full CPU set back afterwards
        old_aff = None; numa = None
        try:
            bus = torch.cuda.get_device_properties(local)
            busid = "%04x:%02x:%02x.0" % (getattr(bus, "pci_domain_id", 0), bus.pci_bus_id, bus.pci_device_id)
            cl = open(f"/sys/bus/pci/devices/{busid}/local_cpulist").read().strip()
            numa = int(open(f"/sys/bus/pci/devices/{busid}/numa_node").read())
            cpus = set()
            for part in cl.split(","):
                a, _, b2 = part.partition("-"); cpus.update(range(int(a), int(b2 or a) + 1))
            if cpus and numa >= 0:
                old_aff = os.sched_getaffinity(0); os.sched_setaffinity(0, cpus & old_aff or old_aff)
        except Exception:
            old_aff = None
        def pinned(nbytes, dtype):
            p = L.jsgpu_host_alloc(int(nbytes))
            if not p:
                return None
            return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(int(nbytes),)).view(dtype)
        lay = bd.layout
        pix_n = sum((int(l.img_x) * int(l.img_y) + 63) // 64 * 64 for l in lay)
        dib_n = sum((int(l.img_x) * int(l.img_y) * 4 + 255) // 256 * 256 for l in lay)
        blk_n = sum((int(l.blk_xmax) * int(l.blk_ymax) + 63) // 64 * 64 for l in lay)
        mcu_n = sum((int(l.mcu_xmax) * int(l.mcu_ymax) + 31) // 32 * 32 for l in lay)
        outs = {"pix_y": pinned(pix_n * 2, np.int16), "pix_cb": pinned(pix_n * 2, np.int16), "pix_cr": pinned(pix_n * 2, np.int16),
                "dib": pinned(dib_n, np.uint8), "blk_y": pinned(blk_n * 2, np.int16), "blk_cb": pinned(blk_n * 2, np.int16),
                "blk_cr": pinned(blk_n * 2, np.int16), "mcu_map": pinned(mcu_n * 4, np.uint32),
                "dht_histo": pinned(nimg * 136 * 4, np.uint32), "stats": pinned(nimg * 16 * 4, np.int32)}
        hbits = pinned(bits.size, np.uint8)
        if hbits is not None and all(v is not None for v in outs.values()):
            hbits[:] = bits
            bd.decode_host(darr, hbits, outs)            # warm-up (allocations are grow-only)
            if world > 1:
                dist.barrier()
            t1 = time.perf_counter()
            for _ in range(args.e2e_steps):
                bd.decode_host(darr, hbits, outs)
            dt = time.perf_counter() - t1
            d2h = sum(v.nbytes for v in outs.values())
            te = torch.tensor([dt], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            e2e = {"value": round(total_px * args.e2e_steps / float(te.item()) / 1e6, 1), "unit": UNIT,
                   "h2d_bytes_per_step": int(bits.size), "d2h_bytes_per_step": int(d2h), "steps": args.e2e_steps,
                   "what": "jsgpu_decode_batch_host: pinned bitstream -> all reference outputs in pinned host memory "
                           "(8 image ranges, D2H of one overlapping upload+decode of the next)"}
            if rank == 0:                                 # the host buffers themselves against the oracle
                ok = True
                for i in checked:
                    ok = ok and not JC.compare(orc.decode(bytes(jpegs[i])), bd.fetch_host(i, outs), what=WHAT)
                e2e["host_buffers_bit_exact_vs_oracle"] = bool(ok)
        else:
            e2e = {"value": None, "unit": UNIT, "error": "pinned host allocation failed"}
        if old_aff is not None:
            os.sched_setaffinity(0, old_aff)
            if e2e is not None: e2e["host_numa_node"] = numa

    # ---- roofline of the dominant kernel + CPU baseline (rank 0) ---------------------------------------
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        bpp = {"420": 13.0, "422": 14.0, "444": 16.0}[ss]
        # stage 2 = the fused dequant+IDCT+upsample+colour kernel (k_idct_tile): algorithmic bytes =
        # coefficient rows read (2 B x samples) + three int16 maps + BGRA written, per padded pixel
        idct_ms = float(stage_ms[2]); huff_ms = float(stage_ms[1])
        ach = bd.npadded_pixels * bpp / (idct_ms / 1e3) / 1e9
        huff_gbs = (bits.size + bd.npadded_pixels * (bpp - 10.0)) / (huff_ms / 1e3) / 1e9      # bitstream read + coefficient rows written
        traffic_rw = None
        traffic = None        # DRAM bytes of one K2 launch from the committed ncu capture of this workload (GB), if there is one
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.config)
            if tj and nimg == CONFIGS[args.config][0]:
                traffic = round(tj["dram_read_gb"] + tj["dram_write_gb"], 3); traffic_rw = (tj["dram_read_gb"], tj["dram_write_gb"])
        except Exception:
            pass
        roof = {"bound": "hbm", "kernel": "k_idct_tile (dequant+IDCT+upsample+YCC->BGRA, stage B)", "achieved": round(ach, 1), "peak": peak,
                "unit": "GB/s", "frac": round(ach / peak, 4), "traffic": traffic, "traffic_unit": "GB per launch (algorithmic: %.3f)" % (bd.npadded_pixels * bpp / 1e9), "peak_source": peak_src,
                "algorithmic_bytes_per_padded_px": bpp, "ms_per_launch": round(idct_ms, 3),
                "stage_ms": {"marker_scan+unstuff": round(float(stage_ms[0]), 3), "huffman": round(huff_ms, 3),
                             "idct+colour": round(idct_ms, 3), "finalize": round(float(stage_ms[3]), 3), "step_total": round(float(stage_ms[4]), 3)},
                "huffman_achieved_gbs": round(huff_gbs, 1)}
        if traffic_rw:        # read-only and write-only rates of the same launch
            roof["dram_read_gbs"] = round(traffic_rw[0] / (idct_ms / 1e3), 1); roof["dram_write_gbs"] = round(traffic_rw[1] / (idct_ms / 1e3), 1)
        cb = None
        if not args.no_cpu and world == 1:          # the CPU leg is reported at N=1 only (the reference arm covers N>1)
            cb, _, _ = cpu_baseline(jpegs, w * h)
        line = {"metric": METRIC, "value": round(value, 1), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(ms_max / args.steps, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "int32 IDCT + f32 colour (int16/u8 outputs)", "data": "synthetic",
                "config": {"workload": f"batch={nimg} {w}x{h} 4:2:0 baseline JPEG per GPU, q{q}, RST interval={ri} MCUs "
                                       f"(BASELINE configs[1]); image-sharded, shared DHT/DQT broadcast over NCCL",
                           "l2": "inputs larger than L2: bitstream %.0f MB + coefficient rows %.0f MB per step" % (bits.size / 1e6, bd.npadded_pixels * 3 / 1e6),
                           "decoder_status_words": status, "bit_exact_vs_oracle": parity_ok, "bit_exact_checked_images": checked,
                           "gen_s": round(t_gen, 1)},
                "roofline": roof, "cpu_baseline": cb, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
