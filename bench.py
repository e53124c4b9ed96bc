#!/usr/bin/env python
"""bench.py — match-finding input MB/s of the B200 engine (and of the reference on the host CPU).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one rank per GPU)
  python bench.py --impl reference [--gpus N] ...                reference arm (rank 0, host CPU)

A step is one pass of the hot path — all four finders (HT2+HT3+BT4 exhaustive+RK256, "R2"
semantics) over one batch of synthetic input:
  N=1  BASELINE.json configs[1]: 100 MB enwik8-shaped text, -window:24 (C2)
  N>1  weak scaling: the input is N x 100 MB, replicated in every GPU's HBM, rank r owns positions
       [r*100 MB, (r+1)*100 MB) and builds its structures over its range plus the 16 MB window
       behind it; no data-path collective. value = all positions / max-over-ranks time.
Prints ONE JSON line (rank 0).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

PER_GPU_BYTES = 100_000_000
HIST_BITS = 24
CPU_SAMPLE = 1 << 23          # 8 MiB keeps -window:24 after the reference's shrink rule (flen >= 2^23)
METRIC = "match-finding input MB/s"


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------------
# reference arm: the reference's own finder objects on the host CPU (oracle/_ref), else the C port
# --------------------------------------------------------------------------------------------------

def cpu_matcher_mbs(x_sample: np.ndarray, hist_bits: int):
    """(MB/s, kind, description): reference match finding (HT2+HT3+BT4+RK256 with the shipped 256-test
    cap, carry + skip rule as in parse_table, no pricing / coding) on one host thread."""
    from oracle import refbind as rb
    if rb.available():
        _, secs = rb.matchfind(x_sample, hist_bits, mode=rb.R0, use_carry=True, dump_mask=0)
        return x_sample.size / secs / 1e6, "reference", secs
    from oracle import oracle as orc
    t = time.perf_counter()
    orc.find(x_sample, hist_bits, orc.F_ALL, 256)
    secs = time.perf_counter() - t
    return x_sample.size / secs / 1e6, "port", secs


COMPRESS_SAMPLE = 16 << 20


def compress_ours(x: np.ndarray, hist_bits: int, device: int):
    """End-to-end compress (BASELINE.json metric, second half) through this repo's own host pipeline
    (libnlzm_codec: parser + model + rANS writer over the engine), host bytes in, stream bytes out."""
    try:
        from nlzm_b200 import codec
        codec.compress(x[:1 << 20], hist_bits, device=device)          # warm-up
        t = time.perf_counter()
        blob, st = codec.compress(x, hist_bits, device=device, with_stats=True)
        secs = time.perf_counter() - t
        t = time.perf_counter()
        ok = codec.decompress(blob) == x.tobytes()
        dsecs = time.perf_counter() - t
        return {"value": x.size / secs / 1e6, "unit": "MB/s", "ratio": len(blob) / x.size, "stream_bytes": len(blob),
                "sample": f"first {x.size} bytes of the same text, -window:{hist_bits}", "roundtrip": bool(ok),
                "decompress_MBps": x.size / dsecs / 1e6, "engine_wait_ms": st["ms_engine_wait"],
                "path": "nlzm_codec_compress: engine blocks double-buffered against one host thread of parsing + rANS coding"}
    except Exception as e:  # the headline matcher numbers must not depend on this leg
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def compress_reference(x: np.ndarray, hist_bits: int):
    """The pristine reference CLI (oracle/_ref/nlzm_r0 -window:N c) on a bounded sample, one thread."""
    try:
        import tempfile
        from oracle import refbind as rb
        if not os.path.exists(rb.REF_R0):
            return {"unavailable": "oracle/_ref/nlzm_r0 not built"}
        with tempfile.TemporaryDirectory() as td:
            src, dst = os.path.join(td, "in.bin"), os.path.join(td, "out.nlzm")
            x.tofile(src)
            t = time.perf_counter()
            rb.r0_cli(f"-window:{hist_bits}", "c", src, dst)
            secs = time.perf_counter() - t
            size = os.path.getsize(dst)
        return {"value": x.size / secs / 1e6, "unit": "MB/s", "ratio": size / x.size, "stream_bytes": size,
                "sample": f"first {x.size} bytes of the same text, -window:{hist_bits}", "cores": 1,
                "path": "nlzm_r0 c (unmodified reference, file in / file out)"}
    except Exception as e:
        return {"unavailable": f"{type(e).__name__}: {e}"[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nlzm_b200 import synth
    xc = synth.text(COMPRESS_SAMPLE, 1)     # same generator and seed as the GPU workload: its first 16 MiB (compress leg:
    x = xc[:CPU_SAMPLE]                     # the same bytes as our arm's compress leg), the matcher sample = its first 8 MiB
    vals, kind = [], "reference"
    for i in range(args.warmup + args.steps):
        mbs, kind, secs = cpu_matcher_mbs(x, HIST_BITS)
        if i >= args.warmup:
            vals.append((mbs, secs))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([s for _, s in vals]) * 1e3)
    sample = f"first {CPU_SAMPLE} bytes of the C2 text (seed 1), -window:{HIST_BITS}, one host thread (the reference has no threads)"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"C2: 100 MB synthetic enwik8-shaped text, -window:{HIST_BITS}; each step = {sample}"},
            "cpu_baseline": {"value": value, "unit": "MB/s", "cores": 1, "kind": kind, "sample": sample,
                             "host_cores_available": os.cpu_count()},
            "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "compress": compress_reference(xc, HIST_BITS)}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------
# our arm
# --------------------------------------------------------------------------------------------------

def _sum_stats(acc, st):
    for k in ("ms_rank", "ms_levels", "ms_cross", "ms_ht", "ms_rk", "ms_merge", "ms_total"):
        acc[k] = acc.get(k, 0.0) + float(getattr(st, k))
    acc["segments_queried"] = acc.get("segments_queried", 0) + int(st.segments_queried)


def run_ours(args):
    import torch
    import torch.distributed as dist
    from nlzm_b200 import synth, sharding
    from nlzm_b200.matchfinder import MatchFinders, MatchFinderError, profile, kernel_times, geometry

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    gloo = None
    if world > 1:
        # NCCL may print its version banner on stdout; the contract is ONE JSON line there
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            gloo = dist.new_group(backend="gloo")          # descriptor exchange of the segment hand-over (host side)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def replicate(make):
        """input generated once on rank 0 and replicated into every GPU's HBM (setup, not the hot path)"""
        if rank == 0:
            x_pin = torch.from_numpy(make()).pin_memory()
            n = torch.tensor([x_pin.numel()], dtype=torch.int64, device=dev)
        else:
            x_pin, n = None, torch.zeros(1, dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(n, src=0)
        x_dev = x_pin.to(dev) if rank == 0 else torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
        if world > 1:
            dist.broadcast(x_dev, src=0)
            if rank != 0:
                x_pin = x_dev.cpu().pin_memory()
        torch.cuda.synchronize()
        return x_pin, x_dev

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # ==============================================================================================
    # main leg: C2 x N (weak scaling). Rank r owns 100 MB of positions; the 16 MB window behind its range
    # comes from rank r-1 as a copy of its sorted blocks over NVLink (no re-ranking of the halo).
    # ==============================================================================================
    total = args.bytes_per_gpu * world
    x_pin, x_dev = replicate(lambda: synth.text(total, 1))
    own_b, own_e = sharding.shard_range(total, rank, world)
    n_own = own_e - own_b
    g = geometry(total, args.hist_bits)
    blocks = sharding.blocks_for(own_b, own_e, g.window, args.block)

    mf = MatchFinders()
    mf.Init(args.hist_bits, (x_dev.data_ptr(), total), device=local, max_range=max(e - b for b, e in blocks))
    sf = sharding.ShardedFind(mf, rank, world, g.window, group=gloo, transport="ipc")
    state = {"handover": "segments (CUDA IPC peer copy)" if world > 1 else "n/a (one rank)"}

    def step_resident(acc=None):
        """one pass over this rank's positions, input resident in HBM, results left in HBM"""
        out = {"steps": 0, "tuples": 0}

        def find(b, e, i):
            v = mf.find_device(b, e, slot=i & 1)
            st = mf.stats()
            out["steps"] += int(v.n_steps)
            out["tuples"] += int(st.tuples_last)
            if acc is not None:
                _sum_stats(acc, st)
        if state["handover"].startswith("halo"):
            for i, (b, e) in enumerate(blocks):
                find(b, e, i)
        else:
            sf.run(blocks, find)
        return out

    # ---- warm-up (untimed): also sizes every buffer; decides the hand-over transport
    ok = 1
    try:
        step_resident()
    except (MatchFinderError, RuntimeError) as ex:          # e.g. CUDA IPC not permitted in this container
        ok = 0
        sys.stderr.write(f"[bench rank {rank}] segment hand-over failed ({ex}); falling back to halo re-ranking\n")
    if world > 1:
        t_ok = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        ok = int(t_ok.item())
    if not ok:
        state["handover"] = "halo (every rank re-ranks the window behind its range)"
        mf.drop_segments()
    for _ in range(max(args.warmup, 3) - 1):
        flush.zero_()
        step_resident()

    # ---- timed region: K resident steps between barriers; device time from the engine's CUDA events beside it
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    profile(True)
    launches0 = int(mf.stats().kernel_launches)
    acc = {}
    t_wall0 = time.perf_counter()
    n_steps = n_tuples = 0
    for _ in range(args.steps):
        flush.zero_()                        # L2 flush between timed iterations (256 MiB > 126 MB L2)
        o = step_resident(acc)
        n_steps, n_tuples = o["steps"], o["tuples"]
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    launches = int(mf.stats().kernel_launches) - launches0
    kt = kernel_times()
    profile(False)
    clocks = sampler.stop() if sampler else None
    dev_ms = acc.get("ms_total", 0.0)

    # ---- end-to-end: K steps with HOST buffers through the public calls: set_input (H2D) + submit/fetch (compute +
    #      D2H into pinned memory). The device->host copy of step k runs on the engine's copy stream while step k+1
    #      uploads and computes; every step's result is read on the host inside the timed region.
    e2e_trace = []

    def step_e2e_all(k_steps):
        d2h, pending = 0, None
        for k in range(k_steps):
            t_a = time.perf_counter()
            rc = mf._L.nlzm_mf_set_input(mf._h, C.c_void_p(x_pin.data_ptr()), total)
            assert rc == 0
            t_b = time.perf_counter()

            def find(b, e, i, k=k):
                mf.submit(b, e, (k + i) & 1)
                return (k + i) & 1
            if state["handover"].startswith("halo"):
                slots = [find(b, e, i) for i, (b, e) in enumerate(blocks)]
            else:
                slots = sf.run(blocks, find)
            t_c = time.perf_counter()
            if pending is not None:
                off, st = mf.fetch(pending, copy=False)
                d2h += off.nbytes + st.nbytes
                _ = int(off[-1]) + (int(st["len"][-1]) if st.size else 0)
            t_d = time.perf_counter()
            sst = mf.stats()
            e2e_trace.append({"set_input_ms": round((t_b - t_a) * 1e3, 1), "submit_ms": round((t_c - t_b) * 1e3, 1),
                              "fetch_prev_ms": round((t_d - t_c) * 1e3, 1), "last_find_ms": round(float(sst.ms_total), 1),
                              "last_d2h_ms": round(float(sst.ms_d2h), 1)})
            if len(slots) == 1:
                pending = slots[0]
            else:                                            # several blocks per step: drain them in order
                for sl in slots:
                    off, st = mf.fetch(sl, copy=False)
                    d2h += off.nbytes + st.nbytes
                pending = None
        if pending is not None:
            off, st = mf.fetch(pending, copy=False)
            d2h += off.nbytes + st.nbytes
            _ = int(off[-1]) + (int(st["len"][-1]) if st.size else 0)
        return d2h
    e2e_ok = len(blocks) <= 2
    if e2e_ok:
        step_e2e_all(2)                                      # both result slots: their pinned buffers are allocated here
    barrier()
    t0 = time.perf_counter()
    d2h_bytes = step_e2e_all(args.steps) if e2e_ok else 0
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3

    t = torch.tensor([dev_ms, wall_ms, e2e_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([n_steps, n_tuples, launches, d2h_bytes], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms, wall_ms, e2e_ms = [float(v) for v in t.tolist()]
    all_steps, all_tuples, all_launches, d2h_bytes = [int(v) for v in cnt.tolist()]
    imported = sf.imported_bytes
    mf.Release()
    del x_dev

    # ==============================================================================================
    # C3 leg (BASELINE.json configs[2]): ONE 1 GB enwik9-shaped file, -window:28, its position range sharded
    # over the N ranks (strong scaling); reported as an extra key, the headline value stays C2.
    # ==============================================================================================
    c3 = None
    if args.c3:
        try:
            c3 = run_c3(args, torch, dist, dev, local, rank, world, gloo, replicate, barrier, flush, state)
        except (MatchFinderError, RuntimeError) as ex:      # raised on every rank at the same protocol point
            c3 = {"unavailable": f"{type(ex).__name__}: {ex}"[:300]}

    if rank == 0:
        ms_per_step = wall_ms / args.steps
        value = total / (ms_per_step * 1e-3) / 1e6
        e2e_value = total / (e2e_ms / args.steps * 1e-3) / 1e6 if e2e_ok else None
        # Roofline. Algorithmic bytes, SURVEY.md §8(d): short/medium stage 13*N + 8*steps, RK256 stage 5.016*N.
        # `frac` is the whole step against the measured HBM peak (the honest figure for an exhaustive matcher that
        # makes ~25 passes where the metric assumes one); the dominant kernel is shown beside it with the bytes one
        # of its launches has to move by construction (one read + one write of the 32-byte level elements).
        peak, peak_src = measured_hbm_peak()
        top_name, (top_launches, top_ms) = max(kt.items(), key=lambda kv: kv[1][1])
        n_rank0 = n_own
        b_short = 13.0 * n_rank0 + 8.0 * (all_steps / world)
        b_step = b_short + 5.016 * n_rank0
        step_dev_ms = acc.get("ms_total", 0.0) / args.steps
        short_ms = (acc.get("ms_rank", 0) + acc.get("ms_levels", 0) + acc.get("ms_cross", 0) + acc.get("ms_ht", 0) +
                    acc.get("ms_merge", 0)) / args.steps
        avg_launch_ms = top_ms / max(top_launches, 1)
        launches_per_step = top_launches / args.steps
        per_launch_bytes = b_short / max(launches_per_step, 1)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(top_name)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": top_name,
                    "achieved": b_step / (step_dev_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                    "frac": b_step / (step_dev_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
                    "scope": "whole step on rank 0: algorithmic bytes of both stages / device time of the step",
                    "algorithmic_bytes_per_step": b_step,
                    "short_stage": {"algorithmic_bytes": b_short, "ms": short_ms,
                                    "frac": b_short / (short_ms * 1e-3) / 1e9 / peak if short_ms else None},
                    "dominant_kernel": {"name": top_name, "avg_launch_ms": avg_launch_ms, "launches_per_step": launches_per_step,
                                        "share_of_step": top_ms / args.steps / step_dev_ms if step_dev_ms else None,
                                        "algorithmic_bytes_per_launch": per_launch_bytes,
                                        "achieved": per_launch_bytes / (avg_launch_ms * 1e-3) / 1e9,
                                        "frac": per_launch_bytes / (avg_launch_ms * 1e-3) / 1e9 / peak,
                                        "rule": "B_short / launches per step (its launches together do the short/medium stage)"},
                    "kernel_ms_per_step": {k: round(v[1] / args.steps, 3) for k, v in sorted(kt.items(), key=lambda kv: -kv[1][1])}}
        # CPU baseline: the reference's own matchers on one host core, bounded sample of the same workload
        cpu_mbs, cpu_kind, cpu_secs = cpu_matcher_mbs(x_pin.numpy()[:CPU_SAMPLE], args.hist_bits)
        line = {"metric": METRIC, "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"C2 x{world}: {total} B synthetic enwik8-shaped text (seed 1), -window:{args.hist_bits} "
                                       f"(hist_bits {g.hist_bits}), finders HT2+HT3+BT4(exhaustive)+RK256, R2 semantics",
                           "positions_per_gpu": n_own, "blocks_per_step": len(blocks),
                           "sharding": "input replicated per GPU, position range sharded, no collective; the window behind a "
                                       "range is a one-sided copy of the owning rank's sorted blocks: " + state["handover"],
                           "handover_bytes_per_step_rank_max": imported,
                           "l2": "256 MiB flush buffer written between timed steps; working set ~5 GB >> 126 MB L2",
                           "timing": "wall clock around the K steps between barrier + synchronize, max over ranks; "
                                     "device_ms_per_step = CUDA events on the engine's stream",
                           "staircase_steps": all_steps, "candidate_tuples": all_tuples},
                "device_ms_per_step": dev_ms / args.steps,
                "host_phase_ms_rank0": {k: round(v, 2) for k, v in sf.ms.items()},
                "stage_ms_per_step_rank0": {k: round(v / args.steps, 3) for k, v in acc.items() if k.startswith("ms_")},
                "e2e": {"value": e2e_value, "unit": "MB/s", "h2d_bytes_per_step": total * world, "d2h_bytes_per_step": d2h_bytes // max(args.steps, 1),
                        "ms_per_step": e2e_ms / args.steps, "trace_rank0": e2e_trace[-args.steps:],
                        "path": "per step: nlzm_mf_set_input(pinned host) + nlzm_mf_submit / nlzm_mf_fetch (host view in pinned "
                                "memory); the copy of step k overlaps the upload + compute of step k+1; wall clock over K steps"},
                "gpu_launches": all_launches,
                "roofline": roofline,
                "cpu_baseline": {"value": cpu_mbs, "unit": "MB/s", "cores": 1, "kind": cpu_kind,
                                 "sample": f"first {CPU_SAMPLE} bytes of the same text, -window:{args.hist_bits}, reference finders "
                                           f"with the shipped 256-test cap + carry/skip rule, {cpu_secs:.1f} s",
                                 "host_cores_available": os.cpu_count()},
                "clocks": clocks}
        if c3 is not None:
            line["c3"] = c3
        if world == 1 and args.compress:
            line["compress"] = compress_ours(x_pin.numpy()[:COMPRESS_SAMPLE], args.hist_bits, local)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c3(args, torch, dist, dev, local, rank, world, gloo, replicate, barrier, flush, state):
    from nlzm_b200 import synth, sharding
    from nlzm_b200.matchfinder import MatchFinders, geometry
    n3, hb3 = args.c3_bytes, 28
    x_pin, x_dev = replicate(lambda: synth.make("text_drift", n3))
    g = geometry(n3, hb3)
    own_b, own_e = sharding.shard_range(n3, rank, world)
    blocks = sharding.blocks_for(own_b, own_e, g.window)
    mf = MatchFinders()
    mf.Init(hb3, (x_dev.data_ptr(), n3), device=local, max_range=max(e - b for b, e in blocks))
    sf = sharding.ShardedFind(mf, rank, world, g.window, group=gloo, transport="ipc")
    halo = state["handover"].startswith("halo")

    def one_step(acc):
        out = {"steps": 0}

        def find(b, e, i):
            v = mf.find_device(b, e, slot=i & 1)
            out["steps"] += int(v.n_steps)
            _sum_stats(acc, mf.stats())
        if halo:
            for i, (b, e) in enumerate(blocks):
                find(b, e, i)
        else:
            sf.run(blocks, find)
        return out["steps"]

    one_step({})                                            # warm-up: sizes the buffers, maps the neighbours' export buffers
    sf.ms = {}
    barrier()
    acc, steps_timed, n_steps = {}, 2, 0
    t0 = time.perf_counter()
    for _ in range(steps_timed):
        flush.zero_()
        n_steps = one_step(acc)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    per_rank = {"rank": rank, "positions": own_e - own_b, "blocks": len(blocks), "staircase_steps": n_steps,
                "handover_bytes": sf.imported_bytes}
    per_rank.update({k: round(v / steps_timed, 2) for k, v in acc.items() if k.startswith("ms_")})
    per_rank["segments_queried"] = acc.get("segments_queried", 0) // steps_timed
    per_rank["host_phase_ms"] = {k: round(v / steps_timed, 2) for k, v in sf.ms.items()}
    lst = mf.stats()
    per_rank["last_import"] = {"bytes": int(lst.bytes_imported),
                               "note": "asynchronous peer copies on the copy stream, overlapped with the HT / RK stages "
                                       "(530-555 GB/s when timed alone, profiles/r02_bench_n4.json)"}
    t = torch.tensor([wall_ms], dtype=torch.float64, device=dev)
    allr = [per_rank]
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allr = [None] * world
        dist.all_gather_object(allr, per_rank, group=gloo)
    mf.Release()
    ms = float(t.item()) / steps_timed
    return {"workload": f"C3: {n3} B synthetic enwik9-shaped text (text_drift, seed 2), -window:28 (hist_bits {g.hist_bits}), "
                        "one file, position range sharded over the ranks, all four finders, R2 semantics",
            "metric": METRIC, "value": n3 / (ms * 1e-3) / 1e6, "unit": "MB/s", "ms_per_step": ms, "n_gpus": world,
            "scaling": "strong", "steps": steps_timed, "warmup": 1,
            "timing": "wall clock between barrier + synchronize, max over ranks; results resident in HBM",
            "handover": state["handover"], "staircase_steps": sum(r["staircase_steps"] for r in allr), "per_rank": allr}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes-per-gpu", type=int, default=PER_GPU_BYTES)
    ap.add_argument("--hist-bits", type=int, default=HIST_BITS)
    ap.add_argument("--block", type=int, default=1 << 28, help="largest engine call (positions)")
    ap.add_argument("--c3", type=int, default=1, help="also run the C3 leg (1 GB, -window:28, positions sharded over the ranks)")
    ap.add_argument("--c3-bytes", type=int, default=1_000_000_000)
    ap.add_argument("--compress", type=int, default=1, help="N=1: also run the end-to-end compress leg")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "ours" and args.gpus > 1 and world == 1:
        # convenience: relaunch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
