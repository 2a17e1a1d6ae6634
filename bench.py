#!/usr/bin/env python
"""bench.py — Baum-Welch residues*states/s (fwd+bwd+counts+M-step) of the B200-native HMM engine.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W [--impl reference]`; one JSON line
on rank 0.  A "step" is ONE EM iteration (E-step over the rank's shard, count all-reduce, M-step) of the
32-state, emission-order-3 ergodic HMM of BASELINE.json configs[2] on 1M x 1 kb synthetic DNA sequences PER GPU
(weak scaling: 8 GPUs == the 8M-sequence config).

  value    : whole-job residues*states/s, inputs resident in HBM, CUDA-event time, max over ranks
  e2e      : same metric through the public host API (HigherOrderHMM.train on HOST buffers): every step copies the
             packed residues host->device from pinned memory and reads the objective + parameters back
             (reported twice: from pinned host memory, and from ordinary pageable memory, which the library stages
             through its own pinned buffers)
  roofline : dominant kernel = the fused E-step kernel k_mma_estep; bound = the FP64 rate.  The kernel issues its
             products as FP64 tensor-core instructions (mma.sync.m8n8k4.f64 = SASS DMMA) — on B200 the DMMA and DFMA
             peaks coincide, so the roofline denominator is the same FP64 peak either way;
             `dmma_pipe_frac` relates the DMMA work alone to the measured DMMA peak.  `traffic` comes from the ncu
             capture listed in profiles/traffic_table.json for this (states, order, kernel).
  cpu_baseline : the C oracle (port of the Java path) on the host cores, bounded sample
  multi_gpu_check (N > 1, outside the timed region): a small sharded EM run through the library's NCCL path against a
             1-rank run over the union of the shards on rank 0.
  --scaling strong: the TOTAL number of sequences is fixed (--total-seqs, default 8M = BASELINE configs[2] literally)
             and divided over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "Baum-Welch residues*states/sec (fwd+bwd+counts)"
UNIT = "residues*states/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--states", type=int, default=32)
    ap.add_argument("--order", type=int, default=3)
    ap.add_argument("--nseq", type=int, default=1_000_000, help="sequences per GPU")
    ap.add_argument("--len", type=int, default=1000)
    ap.add_argument("--fast", type=int, default=1)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seqs", type=int, default=0, help="sequences of the CPU sample (0: 400 per core)")
    ap.add_argument("--set", action="append", default=[], metavar="NAME=VALUE", help="library tuning option (jh_set_option), repeatable")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-seqs", type=int, default=8_000_000, help="--scaling strong: sequences of the whole job")
    ap.add_argument("--no-multi-check", action="store_true")
    args = ap.parse_args()
    if args.scaling == "strong":
        args.nseq = args.total_seqs // max(1, args.gpus)
    return args


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_sample(args, threads: int):
    """Oracle (C port of the Java path) on a bounded sample of the same workload: one E-step."""
    from jstacs_b200.workloads import ergodic_hmm, make_data
    from oracle import binding

    n = args.cpu_seqs or 400 * threads  # ~10-20 s of host work on this workload
    hmm = ergodic_hmm(args.states, args.order, ess=4.0)
    data = make_data(n, args.len)
    om = binding.OracleModel(hmm.ir())
    om.estep(data.codes[:data.offsets[min(threads, n)]], data.offsets[:min(threads, n) + 1], threads=threads)  # warm-up
    t0 = time.perf_counter()
    ts, es, sc = om.estep(data.codes, data.offsets, threads=threads)
    om.mstep(ts, es)
    dt = time.perf_counter() - t0
    return data.getNumberOfResidues() * args.states / dt, dt, n


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  Jstacs is Java and this image has no JVM, so
    the arm times the C port (oracle/) with all host threads on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as g

    g.build()
    threads = os.cpu_count() or 1
    vals, n = [], 0
    for i in range(args.warmup + args.steps):
        v, dt, n = cpu_sample(args, threads)
        if i >= args.warmup:
            vals.append((v, dt))
        if sum(d for _, d in vals) > 120:
            break
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([d for _, d in vals]) * 1e3)
    sample = f"{n} x {args.len} bp sequences per step (of {args.nseq * args.gpus}), one E-step + M-step, C port of HigherOrderHMM.baumWelch"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(vals), "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"BASELINE configs[2]: Baum-Welch, {args.states}-state ergodic HMM, order-{args.order} DNA emissions, "
                        f"{args.nseq * world} x {args.len} bp java.util.Random sequences ({args.nseq} per GPU)",
            "states": args.states, "emission_order": args.order, "n_seq_per_gpu": args.nseq, "seq_len": args.len,
            "parallelism": f"dp{world} (sequences sharded, one NCCL all-reduce of the count block per iteration)",
            "scaling": args.scaling,
            "l2": "inputs larger than L2 (residue buffer >= 1 GB per GPU; forward scratch > 1 GB)", "ess": 4.0}


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
        return
    import torch
    import torch.distributed as dist

    import __graft_entry__ as g

    g.build()
    from jstacs_b200 import _native
    from jstacs_b200.hmm import BaumWelchParameterSet, IterationCondition
    from jstacs_b200.parallel import broadcast_comm_id
    from jstacs_b200.workloads import ergodic_hmm, make_data

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    L = _native.lib()
    _native.check(L.jh_set_device(local))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.current_stream()
    # torch's default stream has handle 0; 0x1 is cudaStreamLegacy, the same stream addressed explicitly
    _native.check(L.jh_set_stream(stream.cuda_stream or 1))
    _native.check(L.jh_set_option(b"fast", args.fast))
    _native.check(L.jh_set_option(b"fast_variant", args.variant))
    for kv in args.set:
        name, value = kv.split("=")
        _native.check(L.jh_set_option(name.encode(), int(value)))

    S = args.states
    hmm = ergodic_hmm(S, args.order, n_iter=1, ess=4.0)
    t_gen = time.perf_counter()
    data = make_data(args.nseq, args.len, start_index=rank * args.nseq)  # rank r owns sequences [r*nseq, (r+1)*nseq)
    t_gen = time.perf_counter() - t_gen
    eng = hmm._engine()
    if world > 1:
        import ctypes as C

        def make_id():
            buf = (C.c_uint8 * 128)()
            _native.check(L.jh_comm_unique_id(buf))
            return bytes(buf)

        cid = broadcast_comm_id(make_id, device=torch.device("cuda", local))
        idb = (C.c_uint8 * 128).from_buffer_copy(cid)
        _native.check(L.jh_comm_init(eng.h, idb, world, rank))
    # pinned host copy of the packed residues (e2e path) and the resident device copy (value path)
    pinned = torch.from_numpy(data.codes).pin_memory()
    codes_host = pinned.numpy()
    nd = _native.NativeData(codes_host, data.offsets, None, 4)
    residues = data.getNumberOfResidues()
    total_units = residues * world * S

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def em_step():  # exactly one E-step (+all-reduce) and one M-step
        eng.estep(nd, fetch=False)
        eng.mstep()

    launches0 = L.jh_kernel_launches()
    for _ in range(args.warmup):
        em_step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches1 = L.jh_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    e0.record(stream)
    for _ in range(args.steps):
        em_step()
        kernel_ms.append(None)
    e1.record(stream)
    barrier()
    launches2 = L.jh_kernel_launches()
    ms_total = e0.elapsed_time(e1)
    # dominant-kernel time: CUDA events around the E-step kernel launches, measured by the library on the same stream
    kms = []
    for _ in range(min(args.steps, 3)):
        eng.estep(nd, fetch=False)
        kms.append(eng.last_main_kernel_ms())
        eng.mstep()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = total_units / (ms_step * 1e-3)

    # ---- e2e: public host API, H2D of the inputs and D2H of the results inside the timed region ----
    e2e = None
    if not args.no_e2e:
        hmm.trainingParameter = BaumWelchParameterSet(1, IterationCondition(2), 1)
        from jstacs_b200.data import DataSet

        def e2e_run(codes):
            def e2e_step():
                ds = _packed(DataSet, codes, data.offsets)
                hmm.train(ds)  # packs -> H2D, E-step, M-step, E-step(objective), reads objective + parameters back
                ds._native.close()

            e2e_step()
            barrier()
            t0 = time.perf_counter()
            n_e2e = max(1, min(args.steps, 2))
            for _ in range(n_e2e):
                e2e_step()
            barrier()
            dt = torch.tensor([(time.perf_counter() - t0) / n_e2e], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            # IterationCondition(2) = 2 E-steps + 1 M-step per train() call: two passes over the data per step
            return 2 * total_units / float(dt.item())

        n_par = sum(eng.sizes)
        e2e_pinned = e2e_run(codes_host)
        e2e_pageable = e2e_run(data.codes)  # an ordinary numpy array: the library stages it through its pinned buffers
        e2e = {"value": e2e_pinned, "unit": UNIT, "h2d_bytes_per_step": int(residues + 8 * (data.offsets.size + n_par)),
               "d2h_bytes_per_step": int(8 * (n_par + 2)), "pinned": True,
               "pageable": {"value": e2e_pageable, "pinned": False, "note": "same call from pageable host memory (library-side pinned staging)"},
               "note": "HigherOrderHMM.train(DataSet) with IterationCondition(2): pack->H2D, 2 E-steps + 1 M-step, objective + parameters read back; wall clock incl. host work"}

    # ---- multi-GPU correctness (outside every timed region): the library's NCCL path against a 1-rank run ----
    multi_check = None
    if world > 1 and not args.no_multi_check:
        multi_check = multi_gpu_check(args, eng, hmm, world, rank, torch, dist, _native)

    if rank == 0:
        import ctypes as C

        peak = None
        try:
            v = C.c_double()
            _native.check(L.jh_measure_fp64_peak(C.byref(v)))
            peak = v.value
        except Exception:
            peak = None
        # executed useful flops of the scaled formulation: forward 2S^2, backward 2S^2, counts 2S^2 (FMA = 2 flop) + O(S)
        flop_per_res = 6.0 * S * S + 12.0 * S
        nominal_per_res = 18.0 * S * S + 4.0 * S  # SURVEY §8d nominal count of the reference formulation (exp/log = 1 flop)
        k_ms = float(np.mean(kms))
        achieved = flop_per_res * residues / (k_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        # dram bytes per residue of this kernel on this shape: from the ncu --set full capture named in profiles/traffic_table.json
        kernel = ("k_mma_estep" if args.variant == 0 and S >= 5 else "k_fast_estep") if args.fast else "k_exact_fb"
        traffic_per_res, traffic_src = None, None
        try:
            for row in json.load(open(os.path.join(ROOT, "profiles", "traffic_table.json")))["rows"]:
                if row["kernel"] == kernel and row["states"] == S and row["emission_order"] == args.order:
                    traffic_per_res, traffic_src = float(row["dram_bytes_per_residue"]), row["source"]
        except Exception:
            pass
        dmma_peak = None
        try:
            v = C.c_double()
            _native.check(L.jh_probe_dmma(8, 256, 4, C.byref(v)))
            dmma_peak = v.value
        except Exception:
            dmma_peak = None
        # the three products of a position are 12 DMMA.8x8x4 per residue and 8 sequences = 12 * 512 flop per residue for S = 32
        dmma_flop_per_res = 6.0 * S * S if kernel == "k_mma_estep" else 0.0
        roofline = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": (achieved / peak) if peak else None,
                    "traffic": traffic_per_res * residues if traffic_per_res else None, "traffic_source": traffic_src,
                    "dmma_pipe_frac": (dmma_flop_per_res * residues / (k_ms * 1e-3) / 1e12 / dmma_peak) if (dmma_peak and dmma_flop_per_res) else None,
                    "dmma_peak": dmma_peak, "kernel": kernel, "kernel_ms": k_ms,
                    "flop_per_residue": flop_per_res, "peak_source": "DFMA-chain microbenchmark measured live on this GPU (MEASURED_PEAKS.json has no fp64 entry)",
                    "nominal_reference_flop_per_residue": nominal_per_res,
                    "nominal_frac": nominal_per_res * residues / (k_ms * 1e-3) / 1e12 / peak if peak else None,
                    "hbm": {"algorithmic_bytes_per_residue": 2.0, "achieved_gbs": 2.0 * residues / (k_ms * 1e-3) / 1e9,
                            "implementation_bytes_per_residue": traffic_per_res, "peak_gbs": hbm_peak,
                            "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks, "e2e": e2e, "multi_gpu_check": multi_check,
            "gpu_launches": int(launches2 - launches1), "roofline": roofline, "data_gen_s": t_gen,
        }
        if not args.no_cpu_baseline and world >= 1:
            threads = os.cpu_count() or 1
            v, dt, n = cpu_sample(args, threads)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{n} x {args.len} bp sequences, one E-step + M-step ({dt:.1f} s), C port of HigherOrderHMM.baumWelch (no JVM in this image)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        _native.check(L.jh_comm_destroy(eng.h))
        dist.destroy_process_group()
    del launches0


def multi_gpu_check(args, eng, hmm, world, rank, torch, dist, _native):
    """A small sharded Baum-Welch run (4 iterations) through jh_baum_welch with the communicator attached, against the same
    run over the UNION of the shards on rank 0 without a communicator: objectives <= 1e-12 relative, parameters <= 1e-10."""
    from jstacs_b200.workloads import ergodic_hmm, make_data, set_closed_form_parameters

    n_small, n_iter = 2048, 4
    set_closed_form_parameters(hmm)  # back to the closed-form starting point (the timed steps and train() moved the parameters)
    shard = make_data(n_small, 200, start_index=rank * n_small)
    nd = _native.NativeData(shard.codes, shard.offsets, None, 4)
    obj = eng.baum_welch(nd, _native.MODE_BAUM_WELCH, _native.TERM_ITERATIONS, n_iter, 0.0)
    par = np.concatenate(eng.get_params())
    nd.close()
    # every rank must hold the same objective history and parameters (replicated M-step)
    mine = torch.from_numpy(np.concatenate([obj, par])).cuda()
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    agree = all(bool(torch.equal(allr[0], t)) for t in allr)
    if rank != 0:
        return None
    ref = ergodic_hmm(args.states, args.order, n_iter=1, ess=4.0)
    union = make_data(n_small * world, 200)
    e2 = ref._engine()
    nd2 = _native.NativeData(union.codes, union.offsets, None, 4)
    obj1 = e2.baum_welch(nd2, _native.MODE_BAUM_WELCH, _native.TERM_ITERATIONS, n_iter, 0.0)
    par1 = np.concatenate(e2.get_params())
    nd2.close()
    fin = np.isfinite(par1)
    o_rel = float(np.max(np.abs(obj - obj1) / np.abs(obj1)))
    p_abs = float(np.max(np.abs(par[fin] - par1[fin])))
    return {"ok": bool(agree and o_rel <= 1e-12 and p_abs <= 1e-10 and np.array_equal(np.isfinite(par), fin)), "ranks_agree": agree,
            "objective_rel": o_rel, "param_abs": p_abs, "iterations": n_iter, "sequences": n_small * world,
            "what": "jh_baum_welch over NCCL on per-rank shards vs one rank over their union"}


def _packed(DataSet, codes, offsets):
    ds = DataSet.__new__(DataSet)
    ds.annotation, ds.codes, ds.offsets, ds._native = "bench", codes, offsets, None
    from jstacs_b200.data import DNA

    ds.con = DNA
    return ds


if __name__ == "__main__":
    main()
