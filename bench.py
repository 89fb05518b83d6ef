#!/usr/bin/env python3
"""bench.py -- headline benchmark of the QGTC hot path on B200 (see DESIGN.md "Measurement").

Metric (BASELINE.json): bit-GEMM integer TOPS, ops := 2*M*K*N per GEMM regardless of bit widths
(the reference's convention, QGTC_device.cu:420-422), on BASELINE.json configs[1]'s headline shape:
binary adjacency 4096 x 4096 times 64-wide features at 1/2/4/8 bits, output requantized to the
same width (2_7c_QGTC_GEMM_INT8.py).

One "step" = one pass over a batch of G independent graphs at each of the four bit widths
(4 batched launches).  G * 2 MiB of packed adjacency exceeds the 126 MB L2, so every launch streams
its adjacency from HBM (timing rule: inputs larger than L2).  `value` is device-resident
throughput; `e2e` is the same metric through the host-buffer C-ABI entry point
(qgtc_bitmm2bit_host: pinned H2D of the packed operands + GEMM + D2H of the packed result).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--engine auto|mma|tcgen05]
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
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

M = K = 4096
N = 64
BITS = (1, 2, 4, 8)
OPS_PER_GEMM = 2.0 * M * K * N


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--engine", default="auto", choices=["auto", "mma", "tcgen05"])
    ap.add_argument("--graphs", type=int, default=96, help="graphs per batch (96 x 2 MiB adjacency > L2)")
    ap.add_argument("--e2e-graphs", type=int, default=96)
    ap.add_argument("--ref-device", default="auto", choices=["auto", "cuda", "cpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-epoch", action="store_true", help="skip the Cluster-GCN epoch block")
    return ap.parse_args()


# --------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def dist_setup(n):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, local, world


def barrier(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world):
    import torch
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------------------------
def make_batch(G, seed):
    """Synthetic packed operands, produced on the device by our own val2bit: G Bernoulli(0.5)
    adjacencies (dense in non-zero tiles: the no-jump regime of Fig. 8a) and per-bit-width features."""
    import torch
    from qgtc_ppopp22_b200 import capi
    gen = torch.Generator(device="cuda").manual_seed(seed)
    wa = capi.packed_words(capi.LAYOUT_ROW, 1, M, K)
    A = torch.empty((G, wa), dtype=torch.int32, device="cuda")
    for g in range(G):
        a = (torch.rand((M, K), device="cuda", generator=gen) < 0.5).float()
        A[g] = capi.val2bit(a, 1, False, False).view(-1)
    X = {}
    for b in BITS:
        wx = capi.packed_words(capi.LAYOUT_COL128, b, K, N)
        xb = torch.empty((G, wx), dtype=torch.int32, device="cuda")
        for g in range(G):
            x = torch.randint(0, 1 << b, (K, N), device="cuda", generator=gen).float()
            xb[g] = capi.val2bit(x, b, True, False).view(-1)
        X[b] = xb
    torch.cuda.synchronize()
    return A, X


def run_ours(args):
    import torch
    from qgtc_ppopp22_b200 import capi
    rank, local, world = dist_setup(args.gpus)
    engine = {"auto": capi.ENGINE_AUTO, "mma": capi.ENGINE_MMA_SYNC, "tcgen05": capi.ENGINE_TCGEN05}[args.engine]
    G = args.graphs
    A, X = make_batch(G, seed=3 + rank)
    outs = {b: torch.empty((G, capi.packed_words(capi.LAYOUT_ROW, b, M, N)), dtype=torch.int32, device="cuda") for b in BITS}

    def step():
        for b in BITS:
            capi.bitMM2Bit_batched(A, X[b], G, M, K, N, 1, b, b, out=outs[b], engine=engine)

    for _ in range(max(args.warmup, 3)):
        step()
    # clocks are sampled from here on; ~0.3 s of untimed load first so that the samples (20 ms apart)
    # describe the loaded state the timed region runs in
    sampler = ClockSampler(local)
    barrier(world)
    if rank == 0:
        sampler.start()
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < 0.3:
        for _ in range(10):
            step()
        torch.cuda.synchronize()
    # One step = the four batched calls, captured once into a CUDA graph (prep + GEMM kernels and the library's
    # stream-ordered image allocations) and replayed: the timed region holds exactly K replays.
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    step_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(step_graph):
        step()
    for _ in range(3):
        step_graph.replay()
    torch.cuda.synchronize()
    barrier(world)
    launches0 = capi.launch_count()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step_graph.replay()
    t1.record()
    torch.cuda.synchronize()
    barrier(world)
    ms_total = max_over_ranks(t0.elapsed_time(t1), world)
    ms_step = ms_total / args.steps
    ops_step = OPS_PER_GEMM * G * len(BITS) * world
    value = ops_step / (ms_step * 1e-3) / 1e12

    # per-call device time of the op (its two kernels) for the roofline: CUDA events on the launching stream
    # around every eager call, same K steps, still inside the clock-sampled window
    evs = []
    for _ in range(args.steps):
        for b in BITS:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.bitMM2Bit_batched(A, X[b], G, M, K, N, 1, b, b, out=outs[b], engine=engine)
            e1.record()
            evs.append((b, e0, e1))
    torch.cuda.synchronize()
    # graph replays do not pass through the library's launch counter; the eager pass launches the same kernels
    launches = capi.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    per_bit = {}
    for b in BITS:
        d = [e0.elapsed_time(e1) for (bb, e0, e1) in evs if bb == b]
        per_bit[b] = sum(d) / len(d)
    # the dominant kernel alone (gemm_tc05_kernel, without prep_b_kernel and the launch gaps of the call): the
    # library brackets each of its launches with CUDA events on the launching stream in a third pass of K steps
    import ctypes
    L = capi.lib()
    L.qgtc_debug_tc05_time.argtypes = [ctypes.c_int]
    L.qgtc_debug_tc05_kernel_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.qgtc_debug_tc05_time(1)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    L.qgtc_debug_tc05_time(0)
    kt = (ctypes.c_float * (args.steps * len(BITS)))()
    nk = L.qgtc_debug_tc05_kernel_times(kt, len(kt))
    per_bit_kernel = {}
    if nk == args.steps * len(BITS):     # call order: bit widths cycle within a step
        for i, b in enumerate(BITS):
            d = [kt[j] for j in range(i, nk, len(BITS))]
            per_bit_kernel[b] = sum(d) / len(d)
    else:                                # another engine ran: the call is a single kernel
        per_bit_kernel = dict(per_bit)
    kern_ms = sum(per_bit_kernel.values()) / len(per_bit_kernel)
    peaks = measured_peaks()
    # the kernel is timed alone (events around each launch): the burst figure applies; the sustained one (1404 TF/s
    # bf16, measured with the SM clock power-capped to ~1.36 GHz) is exceeded by this engine at N = 128
    int8_peak = 2.0 * peaks["bf16_tflops"]
    achieved = OPS_PER_GEMM * G / (kern_ms * 1e-3) / 1e12
    traffic = None
    try:   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this command (profiles/)
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            tj = json.load(f)
        if tj.get("graphs_per_launch") == G:
            traffic = tj["dram_bytes_per_launch"]
    except Exception:
        pass

    # ---- end to end through the host-buffer C-ABI call (pinned host operands in, packed result read back):
    # the same step (all G graphs at the four bit widths), uploads / GEMMs / downloads pipelined by the library
    Ge = min(args.e2e_graphs, G)
    hA = A[:Ge].cpu().pin_memory()
    hX = {b: X[b][:Ge].cpu().pin_memory() for b in BITS}
    hO = {b: torch.empty((Ge, capi.packed_words(capi.LAYOUT_ROW, b, M, N)), dtype=torch.int32).pin_memory() for b in BITS}

    def e2e_step():
        for b in BITS:
            capi.bitMM2Bit_host_batched(hA, hX[b], Ge, M, K, N, 1, b, b, out_host=hO[b], engine=engine)

    e2e_step()
    barrier(world)
    w0 = time.perf_counter()
    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(e2e_steps):
        e2e_step()
    barrier(world)
    e2e_s = max_over_ranks(time.perf_counter() - w0, world) / e2e_steps
    e2e_val = OPS_PER_GEMM * Ge * len(BITS) * world / e2e_s / 1e12
    h2d = sum((hA.numel() + hX[b].numel()) * 4 for b in BITS)
    d2h = sum(hO[b].numel() * 4 for b in BITS)
    e2e_ok = bool(torch.equal(hO[2][0], outs[2][0].cpu()))

    # spot check of the timed outputs against the oracle (not timed): graph 0, 2-bit
    ok = None
    cpu = None
    if rank == 0:
        import numpy as np
        import c_oracle
        want = c_oracle.bitMM2Bit(A[0].cpu().numpy(), X[2][0].cpu().numpy(), M, K, N, 1, 2, 2).reshape(-1)
        ok = bool(np.array_equal(outs[2][0].cpu().numpy().view(np.uint32), want))
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline_port(A[0].cpu().numpy(), {b: X[b][0].cpu().numpy() for b in BITS})

    epoch = None
    if not args.no_epoch:
        try:
            epoch = cluster_gcn_epoch(world, rank)
        except Exception as e:  # the headline metric must still be printed
            epoch = {"error": repr(e)[:300]}

    if rank == 0:
        line = {
            "metric": "bit-GEMM int TOPS (1/2/4/8-bit) vs INT8 peak; Cluster-GCN epoch ms @1-8 GPU",
            "value": round(value, 3), "unit": "TOPS (2*M*K*N per GEMM)", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8 x u8 -> s32 (1-bit adjacency, 1/2/4/8-bit features)", "data": "synthetic",
            "config": {"workload": "BASELINE configs[1]: bit-GEMM 4096x4096x64, 1-bit adjacency x {1,2,4,8}-bit features, "
                                   "batched over %d graphs per launch" % G,
                       "M": M, "K": K, "N": N, "bits": list(BITS), "graphs_per_launch": G, "engine": args.engine,
                       "l2_policy": "inputs larger than L2 (%d MiB packed adjacency per launch)" % (G * 2),
                       "timed_region": "K replays of one CUDA graph holding the step's four batched calls (8 kernels)",
                       "parallelism": "graph batches sharded across ranks, no data-path collective"},
            "per_bit_ms_per_call": {str(b): round(v, 4) for b, v in per_bit.items()},
            "per_bit_tops_per_call": {str(b): round(OPS_PER_GEMM * G / (v * 1e-3) / 1e12, 2) for b, v in per_bit.items()},
            "per_bit_ms_gemm_kernel": {str(b): round(v, 4) for b, v in per_bit_kernel.items()},
            "per_bit_tops_gemm_kernel": {str(b): round(OPS_PER_GEMM * G / (v * 1e-3) / 1e12, 2) for b, v in per_bit_kernel.items()},
            "e2e": {"value": round(e2e_val, 3), "unit": "TOPS (2*M*K*N per GEMM)", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "qgtc_bitmm2bit_host_batched (pinned host operands -> packed result on host; upload/GEMM/download pipelined)",
                    "graphs_per_step": Ge * len(BITS), "matches_device_result": e2e_ok},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": round(int8_peak, 1), "unit": "TFLOP/s",
                         "frac": round(achieved / int8_peak, 4), "traffic": traffic,
                         "note": "integer ops/s of gemm_tc05_kernel alone (CUDA events around each of its launches, mean over the four bit "
                                 "widths; the call also runs prep_b_kernel: see per_bit_ms_per_call); peak = 2 x %s dense "
                                 "bf16 (burst, kernel timed alone) from MEASURED_PEAKS.json, INT8:bf16 = 2:1 on B200; nominal "
                                 "INT8 peak is 4500" % peaks["source"]},
            "cpu_baseline": cpu,
            "cluster_gcn": epoch,
            "clocks": clocks,
            "parity_spot_check": ok,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def cluster_gcn_epoch(world, rank):
    """Second half of BASELINE.json's metric: Cluster-GCN epoch ms on a synthetic ppi-shaped graph
    (56,944 nodes, ~819K edges, 50 feats, hidden 16, 1500 partitions x 20 per batch = 75 batches, 2-bit),
    batches sharded across the ranks; eager ops and one-CUDA-graph-per-epoch replay.  Rank 0 also times the
    torch restatement of the reference's DGL model on the host cores (reported baseline, not a target)."""
    from qgtc_ppopp22_b200 import cluster as C
    from qgtc_ppopp22_b200 import main_qgtc as drv
    g = C.make_synthetic_graph("ppi", 1500)
    out = {"graph": "synthetic ppi-shaped: %d nodes, %d edges, %d feats" % (g.num_nodes, g.num_edges, g.feat.shape[1]),
           "bits": 2, "hidden": 16, "batches": 75, "n_gpus": world}
    base = ["--dataset", "ppi", "--n-epochs", "10", "--bit_width", "2", "--use_QGTC", "--n-classes", "121"]
    a = drv.build_parser().parse_args(base)
    r = drv.run(a, graph=g, quiet=True)
    out["epoch_ms_eager"] = round(r["epoch_ms"], 3)
    a = drv.build_parser().parse_args(base + ["--cuda-graph"])
    r = drv.run(a, graph=g, quiet=True)
    out["epoch_ms_cuda_graph"] = round(r["epoch_ms"], 3)
    a = drv.build_parser().parse_args(base + ["--cuda-graph", "--streams", "16"])
    r = drv.run(a, graph=g, quiet=True)
    out["epoch_ms_cuda_graph_16_streams"] = round(r["epoch_ms"], 3)
    a = drv.build_parser().parse_args(base + ["--cuda-graph", "--streams", "8"] + (["--cpu-baseline"] if rank == 0 else []))
    r = drv.run(a, graph=g, quiet=True)
    out["epoch_ms_cuda_graph_8_streams"] = round(r["epoch_ms"], 3)
    if "cpu_epoch_ms" in r:
        out["cpu_epoch_ms"] = round(r["cpu_epoch_ms"], 3)
        out["cpu_cores"] = r["cpu_cores"]
        out["cpu_kind"] = "torch restatement of the DGL GraphSAGE forward (modules.py:9-45), fp32, CSR aggregation"
    out["reference_published_epoch_ms_sm86"] = 189.016   # README.md:88 (ppi, RTX 3090-class, for context only)
    return out


def cpu_baseline_port(a_words, x_words_by_bit, budget_s=12.0):
    """The oracle's C restatement (AND+popcount over bit planes, OpenMP) on the host cores, on a
    bounded sample: one 4096x4096x64 GEMM per bit width, repeated while the budget lasts."""
    import c_oracle
    t_total, ops = 0.0, 0.0
    reps = 0
    while t_total < budget_s and reps < 50:
        for b in BITS:
            t = time.perf_counter()
            c_oracle.bitMM2Bit(a_words, x_words_by_bit[b], M, K, N, 1, b, b)
            t_total += time.perf_counter() - t
            ops += OPS_PER_GEMM
        reps += 1
    return {"value": round(ops / t_total / 1e12, 4), "unit": "TOPS (2*M*K*N per GEMM)", "cores": c_oracle.num_threads(), "kind": "port",
            "sample": "%d x (one 4096x4096x64 GEMM at each of 1/2/4/8 bits), oracle/oracle.c" % reps}


# --------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm.  The reference implements this path only in CUDA, so its own code path is
    the extension rebuilt for sm_100a from the unmodified sources (oracle/_ref/QGTC_ref), called
    through its public API (QGTC.bitMM2Bit on device tensors) on the same box; when that cannot be
    loaded or --ref-device cpu is given, the oracle's C port is timed on the host cores instead."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    line_common = {
        "impl": "reference", "metric": "bit-GEMM int TOPS (1/2/4/8-bit) vs INT8 peak; Cluster-GCN epoch ms @1-8 GPU",
        "unit": "TOPS (2*M*K*N per GEMM)", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic",
    }
    ref = None
    if args.ref_device in ("auto", "cuda"):
        try:
            import torch
            import build_ref
            if torch.cuda.is_available():
                torch.cuda.set_device(0)
                ref = build_ref.load_variant("QGTC_ref")
        except Exception as e:
            ref = None
            if args.ref_device == "cuda":
                print(json.dumps({"impl": "reference", "unavailable": "reference extension not loadable: %r" % (e,)}))
                return
    if ref is not None:
        import torch
        Gr = 8  # bounded sample of the workload: 8 graphs per bit width per step
        gen = torch.Generator(device="cuda").manual_seed(3)
        devnull = os.open(os.devnull, os.O_WRONLY)
        A, X = [], {b: [] for b in BITS}
        for g in range(Gr):
            A.append(ref.val2bit((torch.rand((M, K), device="cuda", generator=gen) < 0.5).float(), 1, False, False))
            for b in BITS:
                X[b].append(ref.val2bit(torch.randint(0, 1 << b, (K, N), device="cuda", generator=gen).float(), b, True, False))

        def step():
            for b in BITS:
                for g in range(Gr):
                    ref.bitMM2Bit(A[g], X[b][g], M, K, N, 1, b, b)

        for _ in range(max(1, args.warmup)):
            step()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        s = (time.perf_counter() - w0) / args.steps
        os.close(devnull)
        # kernel-only figure of the reference: its own bitMM2Bit_profile (200 back-to-back launches, prints TFLOPs)
        kernel_only = {}
        try:
            import re
            import tempfile
            for b in BITS:
                with tempfile.TemporaryFile(mode="w+b") as tf:
                    sys.stdout.flush()
                    saved = os.dup(1)
                    os.dup2(tf.fileno(), 1)
                    try:
                        ref.bitMM2Bit_profile(A[0], X[b][0], M, K, N, 1, b, b)
                        torch.cuda.synchronize()
                    finally:
                        os.dup2(saved, 1)
                        os.close(saved)
                    tf.seek(0)
                    m = re.search(r"TFLOPs: ([0-9.]+)", tf.read().decode(errors="ignore"))
                    if m:
                        kernel_only[str(b)] = float(m.group(1))
        except Exception as e:
            kernel_only = {"error": repr(e)[:200]}
        val = OPS_PER_GEMM * Gr * len(BITS) / s / 1e12
        line = dict(line_common)
        line.update({"value": round(val, 3), "ms_per_step": round(s * 1e3, 3),
                     "dtype": "b1 AND+popc (wmma m8n8k128; emulated with IMMA on sm_100a)",
                     "config": {"workload": "BASELINE configs[1]: bit-GEMM 4096x4096x64, %d graphs x {1,2,4,8} bits per step, "
                                            "reference QGTC.bitMM2Bit (unmodified sources rebuilt for sm_100a) on cuda:0" % Gr},
                     "cpu_baseline": {"value": round(val, 3), "unit": "TOPS (2*M*K*N per GEMM)", "cores": 0, "kind": "reference",
                                      "sample": "reference CUDA kernels on the same B200 (the reference has no CPU implementation of this path)"},
                     "reference_kernel_only_tops": kernel_only,
                     "e2e": {"value": round(val, 3), "unit": "TOPS (2*M*K*N per GEMM)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
        print(json.dumps(line))
        return
    # CPU port
    import c_oracle
    rng = np.random.default_rng(3)
    a = c_oracle.val2bit((rng.random((M, K)) < 0.5).astype(np.float32), 1, False, False)
    xs = {b: c_oracle.val2bit(rng.integers(0, 1 << b, (K, N)).astype(np.float32), b, True, False) for b in BITS}

    def step():
        for b in BITS:
            c_oracle.bitMM2Bit(a, xs[b], M, K, N, 1, b, b)

    for _ in range(max(1, min(args.warmup, 2))):
        step()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    s = (time.perf_counter() - w0) / args.steps
    val = OPS_PER_GEMM * len(BITS) / s / 1e12
    line = dict(line_common)
    line.update({"value": round(val, 4), "ms_per_step": round(s * 1e3, 3), "dtype": "b1 AND+popcount (uint64)",
                 "config": {"workload": "BASELINE configs[1]: bit-GEMM 4096x4096x64, one graph x {1,2,4,8} bits per step, oracle C port on host cores"},
                 "cpu_baseline": {"value": round(val, 4), "unit": "TOPS (2*M*K*N per GEMM)", "cores": c_oracle.num_threads(), "kind": "port",
                                  "sample": "one 4096x4096x64 GEMM per bit width per step"},
                 "e2e": {"value": round(val, 4), "unit": "TOPS (2*M*K*N per GEMM)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line))


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
