#!/usr/bin/env python
"""bench.py -- alignment DP cells/s and audio-hours/s of the B200 forced-alignment decoder.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|c4j] [--impl reference]

One "step" = one pass of the whole hot path (emission -> DP -> backtrace/intervals -> compact
results on the host) over one ragged batch of synthetic logits.  The default workload is
BASELINE.json configs[1] ("batch of 256 synthetic utterances, 5-30 s, 20-150 phonemes, on 1 B200",
V = 63); with N > 1 every rank aligns its own batch of the same shape (weak scaling, utterances
sharded by rank, no collective on the data path) and `value` is all ranks' cells over the max time.

`value`      device-resident logits -> results in pinned host memory, CUDA-event timed.
`e2e`        the same batch from pinned HOST logits through hubertfa_b200.pipeline.HostBatchAligner:
             collation (hfa_plan_create), chunked H2D of the logits overlapped with the kernels,
             D2H of the results.
`roofline`   the DP forward stage (the dominant kernels): algorithmic bytes / its CUDA-event time.
`cpu_baseline` / `--impl reference`  the C port of the reference decoder (oracle/hfa_oracle.c; the
             reference itself is Python + numba and cannot travel to the GPU box) on all host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_SETS = 4  # rotating input sets: consecutive steps never touch the same memory (total >> 126 MB L2)

WORKLOADS = {
    #        B     V   (min_s, max_s) (s_lo, s_hi)  description
    "c1": (1, 63, None, None, "configs[0]: single 10 s utterance, T=500, S=40, V=63"),
    "c2": (256, 63, (5, 30), (20, 150), "configs[1]: 256 utterances, 5-30 s, 20-150 phonemes, V=63"),
    "c3": (1, 63, None, None, "configs[2]: one 10 min utterance, T=30000, S=2000, V=63"),
    "c4": (4096, 74, (5, 30), (20, 150), "configs[3]: 4096 mixed-length utterances, jyutping V=74"),
    "c4j": (4096, 39, (5, 30), (20, 150), "configs[3]: 4096 mixed-length utterances, japanese V=39"),
    # sizes between configs[1] and configs[3] (routing threshold experiments)
    "m512": (512, 63, (5, 30), (20, 150), "512 utterances, 5-30 s, 20-150 phonemes, V=63"),
    "m1024": (1024, 63, (5, 30), (20, 150), "1024 utterances, 5-30 s, 20-150 phonemes, V=63"),
    "m2048": (2048, 63, (5, 30), (20, 150), "2048 utterances, 5-30 s, 20-150 phonemes, V=63"),
}


def workload_shapes(name: str, seed: int):
    from hubertfa_b200 import synth
    B, V, dur, srange, desc = WORKLOADS[name]
    if name == "c1":
        T, S = np.array([500], np.int32), np.array([40], np.int32)
    elif name == "c3":
        T, S = np.array([30000], np.int32), np.array([2000], np.int32)
    else:
        T, S = synth.sample_shapes(B, seed=seed, min_s=dur[0], max_s=dur[1], s_lo=srange[0], s_hi=srange[1])
        # length-bucketed collation (north_star item 4): the batch is packed longest utterance first,
        # the way a length-aware data loader hands it over; both arms see the same batch
        order = np.argsort(-T.astype(np.int64), kind="stable")
        T, S = np.ascontiguousarray(T[order]), np.ascontiguousarray(S[order])
    return T, S, V, desc


def make_head(T, V, seed):
    """Synthetic network-head output [sum T, V+2] f32 on the CPU (col 0 edge, col 1 ctc blank,
    cols 2.. frame logits, networks/task/forced_alignment.py:288-291), SURVEY.md 8d distributions."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    head = torch.randn(int(np.sum(T)), V + 2, generator=g, dtype=torch.float32)
    head[:, 2:] *= 3.0
    head[:, 0] *= 2.0
    return head


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def _once(self):
        nv = self.nv
        try:
            self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
            r = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20,
                     "hw_thermal_slowdown": 0x40, "hw_power_brake": 0x80, "sync_boost": 0x10,
                     "applications_clocks": 0x2}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def _loop(self):
        while not self._stop.is_set():
            self._once()
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self):
        if self._thr is not None:
            self._stop.set()
            self._thr.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def cpu_arm(T, S, V, ids_cat, head_np, frame_length, budget_s: float, threads: int = 0):
    """Times the C port of the reference decoder on the host cores; returns (cells/s, sec, reps)."""
    from oracle import c_oracle as oc
    cells = int((T.astype(np.int64) * S).sum())
    nthr = threads or oc.max_threads()
    run = lambda: oc.align_batch(T, S, V, head_np[:, 2:], head_np[:, 0], ids_cat, frame_length, nthr)
    run()  # warm (page faults, thread creation)
    reps, t0 = 0, time.perf_counter()
    while True:
        out = run()
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or reps >= 1000:
            break
    assert out["bad"] == 0
    return cells * reps / dt, dt, reps, nthr


def run_reference(args, rank, world):
    """--impl reference: the CPU implementation of the path on this box's host cores."""
    if rank != 0:
        return
    from hubertfa_b200 import synth
    from oracle import c_oracle as oc
    T, S, V, desc = workload_shapes(args.workload, synth.SEED0)
    ids_list = synth.make_ids_batch(T, S, V, seed=synth.SEED0)
    ids_cat = np.concatenate(ids_list)
    head = make_head(T, V, synth.SEED0).numpy()
    cells = int((T.astype(np.int64) * S).sum())
    frames = int(T.sum())
    nthr = oc.max_threads()
    run = lambda: oc.align_batch(T, S, V, head[:, 2:], head[:, 0], ids_cat, synth.FRAME_SECONDS, nthr)
    # bound the whole run to a few minutes: cap the number of steps by a time budget
    for _ in range(args.warmup):
        run()
    times = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = run()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > 120:
            break
    assert out["bad"] == 0
    k = len(times)
    sec = float(np.sum(times))
    val = cells * k / sec
    line = {
        "impl": "reference", "metric": "dp_cells_per_s", "value": val, "unit": "cells/s",
        "n_gpus": args.gpus, "steps": k, "warmup": args.warmup, "ms_per_step": 1e3 * sec / k,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "audio_hours_per_s": frames * synth.FRAME_SECONDS / 3600 * k / sec,
        "config": {"workload": f"{args.workload}: {desc}", "utterances": int(len(T)), "cells": cells,
                   "frames": frames, "frame_seconds": synth.FRAME_SECONDS},
        "cpu_baseline": {"value": val, "unit": "cells/s", "cores": nthr, "kind": "port",
                         "sample": f"{k} passes over the full {args.workload} batch, C port of "
                                   "tools/alignment_decoder.py (oracle/hfa_oracle.c), one utterance per thread"},
        "e2e": {"value": val, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the C4-sized roofline pass")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--e2e-chunks", type=int, default=0,
                    help="equal-sized upload/compute chunks of the e2e leg (0: the library's default shares)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from hubertfa_b200 import ops, synth, _lib

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: hubertfa_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    try:        # pin this rank to the CPUs next to its GPU (matters for the host-driven e2e leg at N > 1)
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
    except Exception:
        pass
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"      # keep NCCL's version banner off stdout: one JSON line only
        dist.init_process_group("nccl", device_id=dev)

    def measure(workload: str, steps: int, warmup: int, do_e2e: bool):
        seed = synth.SEED0 + 100003 * rank           # every rank aligns its own utterances
        T, S, V, desc = workload_shapes(workload, seed)
        ids_list = synth.make_ids_batch(T, S, V, seed=seed)
        ids_cat = np.concatenate(ids_list)
        n_sets = N_SETS if len(T) <= 512 else 2
        heads_host = [make_head(T, V, seed + 17 * i).pin_memory() for i in range(n_sets)]
        heads_dev = [h.to(dev) for h in heads_host]
        plan = ops.AlignPlan(T, S, ids_cat, V, synth.FRAME_SECONDS)
        wss = [plan.new_workspace(dev) for _ in range(n_sets)]
        ress = [plan.new_result(dev) for _ in range(n_sets)]
        host_res = [torch.empty(plan.result_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(n_sets)]
        row0 = plan.frame_off[:-1]

        def set_inputs(p, ws, head):
            base = head.data_ptr()
            st = head.stride(0)
            p.set_inputs(ws, base + (row0 * st + 2) * 4, np.full(p.n_utt, st), np.ones(p.n_utt),
                         base + row0 * st * 4, np.full(p.n_utt, st))

        for ws, head in zip(wss, heads_dev):
            plan.upload(ws)
            set_inputs(plan, ws, head)
        torch.cuda.synchronize()
        dt = _lib.DTYPE_F32
        # small batches: stages 1+2 run as ONE fused pass (emissions computed inside the DP kernel);
        # this is the route hfa_align_batch / decode_batch take on their own
        rt0 = plan.routing()
        unsplit = rt0["warp_utts"] == 0 and rt0["cta_utts"] == 0 and \
            rt0["band_warps"] + rt0["big_band_warps"] == int(len(T))        # hfa_align_batch's own rule
        fused = os.environ.get("HFA_FUSED", "1" if unsplit else "0") != "0"
        if fused:
            try:
                ops.forward_fused(wss[0], plan.handle, dt)
                torch.cuda.synchronize()
            except _lib.HfaError:
                fused = False

        d2h_stream = torch.cuda.Stream()
        d2h_done = [None] * n_sets

        def step(i, evs=None):
            k = i % n_sets
            ws, res = wss[k], ress[k]
            if d2h_done[k] is not None:          # the last download out of this result buffer
                torch.cuda.current_stream().wait_event(d2h_done[k])
            if evs is not None:
                evs[0].record()
            if fused:
                if evs is not None:
                    evs[1].record()
                ops.forward_fused(ws, plan.handle, dt)
            else:
                ops.emission(ws, plan.handle, dt)
                if evs is not None:
                    evs[1].record()
                ops.viterbi_forward(ws, plan.handle, None)
            if evs is not None:
                evs[2].record()
            ops.backtrace(ws, plan.handle, res, None, None)
            if evs is not None:
                evs[3].record()
            # results go to pinned host memory on their own stream: the download of step i overlaps
            # the kernels of step i+1 (every step's results still land on the host inside the region)
            ready = torch.cuda.Event()
            ready.record()
            with torch.cuda.stream(d2h_stream):
                d2h_stream.wait_event(ready)
                host_res[k].copy_(res, non_blocking=True)
                d2h_done[k] = torch.cuda.Event()
                d2h_done[k].record()

        for i in range(warmup):
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        stage_evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(steps)]
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        clocks = ClockSampler(local_rank)
        launches0 = ops.launch_count()
        clocks.start()
        torch.cuda.synchronize()
        e0.record()
        for i in range(steps):
            step(i, stage_evs[i])
        torch.cuda.current_stream().wait_stream(d2h_stream)      # the last download is inside the region
        e1.record()
        if clocks.nv is not None:
            clocks._once()           # the queue is still draining: at least one sample under load even for tiny K
        torch.cuda.synchronize()
        clk = clocks.stop()
        launches = ops.launch_count() - launches0
        ms = e0.elapsed_time(e1)
        st_ms = np.array([[ev[j].elapsed_time(ev[j + 1]) for j in range(3)] for ev in stage_evs]).mean(axis=0)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
            dist.barrier()
        cells = plan.total_cells
        frames = plan.total_frames
        # sanity: results of the last step are real alignments
        v = plan.views(host_res[(steps - 1) % n_sets].numpy())
        assert (v["status"] == 0).all() and (v["n_seg"] > 0).all()

        e2e = None
        if do_e2e:
            from hubertfa_b200.pipeline import BufferPool, HostBatchAligner
            pool = BufferPool(dev)

            def e2e_step(i):
                # collation (4 plans, longest utterances first) + chunked H2D overlapped with the
                # kernels + D2H of the compact results, from pinned host logits
                k = i % n_sets
                al = HostBatchAligner(T, S, ids_cat, V, synth.FRAME_SECONDS, V + 2,
                                      n_chunks=args.e2e_chunks or None, device=dev, pool=pool)
                out = al.run(heads_host[k])
                if not out.all_ok():             # the step's results are read on the host, every step
                    raise RuntimeError("e2e: an utterance was not aligned")
                return out

            for i in range(3):
                e2e_step(i)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for i in range(steps):
                e2e_step(i)
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                e2e_s = float(t.item())
            e2e = {"value": None, "unit": "cells/s", "seconds": e2e_s,
                   "h2d_bytes_per_step": int(heads_host[0].numel() * 4),
                   "d2h_bytes_per_step": int(plan.result_bytes), "ms_per_step": 1e3 * e2e_s / steps}
        alg = plan.algorithmic_bytes(dt)
        if fused:
            alg = dict(alg, emission=0, dp=plan.algorithmic_bytes_fused(dt))
        return dict(fused=fused, T=T, S=S, V=V, desc=desc, ids_cat=ids_cat, head0=heads_host[0], ms=ms, st_ms=st_ms,
                    routing=plan.routing(),
                    cells=cells, frames=frames, clk=clk, launches=launches, e2e=e2e, alg=alg,
                    n_sets=n_sets, bytes_per_set=int(heads_host[0].numel() * 4 + plan.workspace_bytes))

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        hbm_peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json (measured copy)"
    else:
        hbm_peak, peak_src = 6650.0, "fallback of B200_PROFILING.md"
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tpath):
        traffic = json.load(open(tpath))

    def roofline(m, wl):
        dp_s = m["st_ms"][1] * 1e-3
        ach = m["alg"]["dp"] / dp_s / 1e9
        tr = None
        if traffic and wl in traffic:
            tr = traffic[wl].get("dp_dram_bytes_per_step")
        rt = m["routing"]
        names = []
        if rt["band_warps"]:
            names.append(f"hfa_dp_band_kernel<{rt['band_k']}> ({rt['band_warps']} compute warps, several per utterance)")
        if rt["big_band_warps"]:
            names.append(f"hfa_dp_band_kernel<{rt['big_band_k']}> ({rt['big_band_warps']} compute warps, S > 256)")
        if rt["warp_utts"]:
            names.append(f"hfa_dp_warp_any_kernel ({rt['warp_utts']} utterances, one warp each)")
        if rt["cta_utts"]:
            names.append(f"hfa_dp_cta_kernel ({rt['cta_utts']} utterances)")
        em_ms = float(m["st_ms"][0])
        return {"bound": "hbm", "kernel": " + ".join(names) + (
                    " with the emissions computed by its producer warps (fused stages 1+2; + hfa_edge_kernel)"
                    if m["fused"] else " = the DP forward stage of one step"),
                "keeps_dp": rt["keeps_dp"], "fused_emission": m["fused"],
                "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": tr,
                "peak_source": peak_src, "algorithmic_bytes_per_step": m["alg"]["dp"],
                "stage_ms": {"emission": float(m["st_ms"][0]), "dp": float(m["st_ms"][1]),
                             "backtrace": float(m["st_ms"][2])},
                "stage_gbs": {"emission": (m["alg"]["emission"] / (em_ms * 1e-3) / 1e9) if em_ms > 1e-4 else None,
                              "dp": ach, "backtrace": m["alg"]["backtrace"] / (m["st_ms"][2] * 1e-3) / 1e9}}

    m = measure(args.workload, args.steps, args.warmup, do_e2e=True)
    sec = m["ms"] * 1e-3
    total_cells = m["cells"] * world     # every rank has the same shape distribution; exact sum below
    if world > 1:
        t = torch.tensor([m["cells"], m["frames"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_cells, total_frames = float(t[0].item()), float(t[1].item())
    else:
        total_frames = m["frames"]
    m["e2e"]["value"] = total_cells * args.steps / m["e2e"].pop("seconds")
    value = total_cells * args.steps / sec

    line = {
        "metric": "dp_cells_per_s", "value": value, "unit": "cells/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "audio_hours_per_s": total_frames * synth.FRAME_SECONDS / 3600 * args.steps / sec,
        "config": {"workload": f"{args.workload}: {m['desc']}", "utterances_per_gpu": int(len(m["T"])),
                   "cells_per_gpu": int(m["cells"]), "frames_per_gpu": int(m["frames"]),
                   "frame_seconds": synth.FRAME_SECONDS, "sharding": "utterances by rank, no collective",
                   "collation": "batch packed longest utterance first",
                   "e2e_pipeline": "upload cut by rows (30/28/22/12/8 % of the bytes): the DMA of piece i+1 overlaps "
                                   "collation and kernels of the utterances completed by piece i",
                   "l2": f"{m['n_sets']} rotating input+workspace sets of {m['bytes_per_set'] / 1e6:.0f} MB "
                         "(consecutive steps touch different memory; total > 126 MB L2)"},
        "clocks": m["clk"], "e2e": m["e2e"], "gpu_launches": int(m["launches"]),
        "roofline": roofline(m, args.workload),
    }

    if rank == 0 and world == 1 and not args.no_cpu:
        v, dt_s, reps, nthr = cpu_arm(m["T"], m["S"], m["V"], m["ids_cat"], m["head0"].numpy(),
                                      synth.FRAME_SECONDS, args.cpu_seconds)
        line["cpu_baseline"] = {"value": v, "unit": "cells/s", "cores": nthr, "kind": "port",
                                "sample": f"{reps} passes over the full {args.workload} batch in {dt_s:.1f} s, "
                                          "C port of tools/alignment_decoder.py (oracle/hfa_oracle.c), "
                                          "one utterance per thread"}
    else:
        line["cpu_baseline"] = None

    if rank == 0 and world == 1 and not args.no_extra and args.workload == "c2":
        # the machine-filling configuration (configs[3]) for the roofline: same kernels, 16x the batch
        del m
        torch.cuda.empty_cache()
        x = measure("c4", 10, 3, do_e2e=False)
        xs = x["ms"] * 1e-3
        line["extra"] = {"c4": {"workload": f"c4: {x['desc']}", "value": x["cells"] * 10 / xs,
                                "unit": "cells/s", "ms_per_step": x["ms"] / 10,
                                "audio_hours_per_s": x["frames"] * synth.FRAME_SECONDS / 3600 * 10 / xs,
                                "roofline": roofline(x, "c4"), "clocks": x["clk"]}}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
