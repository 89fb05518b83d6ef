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
    ap.add_argument("--passes", type=int, default=16, help="passes over the batch per step (so that K steps are >= 100 ms of device time)")
    ap.add_argument("--e2e-graphs", type=int, default=96)
    ap.add_argument("--ref-device", default="auto", choices=["auto", "cuda", "cpu"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-epoch", action="store_true", help="skip the Cluster-GCN epoch block")
    ap.add_argument("--no-zero-tile", action="store_true", help="skip the zero-tile-jumping / adjacency-size block (BASELINE configs[4])")
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


NUMA_NODE = None     # NUMA node this rank was bound to (the one its GPU hangs off), for the e2e record


def bind_to_gpu_numa_node(local):
    """Pin this process (and so the pinned host buffers it allocates: first touch) to the CPUs of the NUMA node its GPU is attached to.
    With one rank per GPU the end-to-end leg is bound by host memory and the PCIe roots; buffers on the far socket halve it."""
    global NUMA_NODE
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return
        cpus = set()
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            NUMA_NODE = node
    except Exception:
        pass


def dist_setup(n):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if os.environ.get("QGTC_BENCH_NUMA", "1") != "0":
        bind_to_gpu_numa_node(local)
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


def event_timed(fn, reps):
    import torch
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def graph_of(fn):
    """Capture fn() once into a CUDA graph (after one eager call on a side stream) and return the graph."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    return g


def gather_floats(x, world):
    import torch
    if world == 1:
        return [x]
    import torch.distributed as dist
    out = [torch.zeros(1, dtype=torch.float64, device="cuda") for _ in range(world)]
    dist.all_gather(out, torch.tensor([x], dtype=torch.float64, device="cuda"))
    return [float(t.item()) for t in out]


def stats(xs):
    xs = sorted(xs)
    return {"min": round(xs[0], 4), "median": round(xs[len(xs) // 2], 4), "max": round(xs[-1], 4)}


WORKLOAD = ("BASELINE configs[1]: bit-GEMM 4096x4096x64, 1-bit adjacency x {1,2,4,8}-bit features, batches of %d graphs "
            "(%d MiB of packed adjacency per bit width: larger than L2)")


def config_dict(G):
    """Identical for both arms (the reference arm processes a bounded sample of it per step, stated in its cpu_baseline.sample)."""
    return {"workload": WORKLOAD % (G, G * 2), "M": M, "K": K, "N": N, "bits": list(BITS), "graphs_per_batch": G,
            "l2_policy": "inputs larger than L2 (%d MiB packed adjacency per bit width)" % (G * 2),
            "parallelism": "graph batches sharded across ranks, no data-path collective"}


def run_ours(args):
    import torch
    from qgtc_ppopp22_b200 import capi
    rank, local, world = dist_setup(args.gpus)
    engine = {"auto": capi.ENGINE_AUTO, "mma": capi.ENGINE_MMA_SYNC, "tcgen05": capi.ENGINE_TCGEN05}[args.engine]
    G, P = args.graphs, args.passes
    A, X = make_batch(G, seed=3 + rank)
    outs = {b: torch.empty((G, capi.packed_words(capi.LAYOUT_ROW, b, M, N)), dtype=torch.int32, device="cuda") for b in BITS}

    def one_pass():
        for b in BITS:
            capi.bitMM2Bit_batched(A, X[b], G, M, K, N, 1, b, b, out=outs[b], engine=engine)

    def step():          # a step = P passes over the batch at each of the four bit widths (K steps >= 100 ms of device time)
        for _ in range(P):
            one_pass()

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
        step()
        torch.cuda.synchronize()
    # One step captured once into a CUDA graph (8 kernels per pass: prep + GEMM per bit width, plus the library's
    # stream-ordered image allocations) and replayed: the timed region holds exactly K replays.
    step_graph = graph_of(step)
    for _ in range(2):
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
    my_ms = t0.elapsed_time(t1)
    rank_ms = gather_floats(my_ms, world)
    ms_total = max(rank_ms)
    ms_step = ms_total / args.steps
    ops_step = OPS_PER_GEMM * G * len(BITS) * P * world
    value = ops_step / (ms_step * 1e-3) / 1e12

    # per-call device time of the op (its two kernels): CUDA events on the launching stream around every eager call of
    # one pass per step, still inside the clock-sampled window
    evs = []
    for _ in range(args.steps):
        for b in BITS:
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            capi.bitMM2Bit_batched(A, X[b], G, M, K, N, 1, b, b, out=outs[b], engine=engine)
            e1.record()
            evs.append((b, e0, e1))
    torch.cuda.synchronize()
    # graph replays do not pass through the library's launch counter; an eager step launches the same kernels
    step()
    torch.cuda.synchronize()
    launches = (capi.launch_count() - launches0 - 2 * len(BITS) * args.steps) * args.steps
    per_bit = {}
    for b in BITS:
        d = [e0.elapsed_time(e1) for (bb, e0, e1) in evs if bb == b]
        per_bit[b] = sum(d) / len(d)
    # the dominant kernel alone (gemm_tc05_kernel, without prep_b_kernel and the launch gaps of the call): the
    # library brackets each of its launches with CUDA events on the launching stream in a third pass
    import ctypes
    L = capi.lib()
    L.qgtc_debug_tc05_time.argtypes = [ctypes.c_int]
    L.qgtc_debug_tc05_kernel_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.qgtc_debug_tc05_time(1)
    for _ in range(args.steps):
        one_pass()
    torch.cuda.synchronize()
    L.qgtc_debug_tc05_time(0)
    kt = (ctypes.c_float * (args.steps * len(BITS)))()
    nk = L.qgtc_debug_tc05_kernel_times(kt, len(kt))
    per_bit_kernel = {}
    if nk == args.steps * len(BITS):     # call order: bit widths cycle within a pass
        for i, b in enumerate(BITS):
            d = [kt[j] for j in range(i, nk, len(BITS))]
            per_bit_kernel[b] = sum(d) / len(d)
    else:                                # another engine ran: the call is a single kernel
        per_bit_kernel = dict(per_bit)
    clocks = sampler.stop() if rank == 0 else None

    peaks = measured_peaks()
    # Peaks (MEASURED_PEAKS.json, kernel timed alone -> burst figure): INT8 = 2 x bf16 and FP4 = 4 x bf16 on B200.  The 1- and
    # 2-bit launches run the 4-bit operand path (kind::mxf4, E2M1 x E2M1 -> f32, exact below 2^23), the 4- and 8-bit launches
    # the u8 path (kind::i8 -> s32): each path is quoted against the peak of the instruction it uses, and the FP4 launches
    # also against the INT8-equivalent peak the metric names.
    int8_peak, fp4_peak = 2.0 * peaks["bf16_tflops"], 4.0 * peaks["bf16_tflops"]
    tops_k = {b: OPS_PER_GEMM * G / (per_bit_kernel[b] * 1e-3) / 1e12 for b in BITS}
    fp4_bits = [b for b in BITS if b <= 2 and args.engine != "mma" and os.environ.get("QGTC_TC05_FP4", "1") != "0"]
    u8_bits = [b for b in BITS if b not in fp4_bits]
    kern_ms = sum(per_bit_kernel.values()) / len(per_bit_kernel)
    achieved = OPS_PER_GEMM * G / (kern_ms * 1e-3) / 1e12
    traffic = None
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this command (profiles/)
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tj = json.load(f)
            if tj.get("graphs_per_launch") == G:
                traffic = tj["dram_bytes_per_launch"]
                break
        except Exception:
            pass
    by_path = {}
    if u8_bits:
        a_u8 = sum(tops_k[b] for b in u8_bits) / len(u8_bits)
        by_path["u8_kind_i8"] = {"bits": u8_bits, "achieved": round(a_u8, 1), "peak": round(int8_peak, 1), "frac": round(a_u8 / int8_peak, 4),
                                 "unit": "TOPS", "limiter": "three-stage TMEM ring: adjacency fill (CUDA-core bit -> byte expansion, ALU pipe 50-59 %) + MMA pick-up + drain per stage; tensor pipe 46-56 % active (ncu profiles/r2_tc05_ncu_summary_call31.txt, clock64 trace profiles/r2_tc05_trace_u8_n64_call27.txt); i8 dispatch floor 32 cycles / MMA at N = 64 (profiles/r2_mma_dispatch_rate_call20.txt)"}
    if fp4_bits:
        a_f4 = sum(tops_k[b] for b in fp4_bits) / len(fp4_bits)
        by_path["fp4_kind_mxf4"] = {"bits": fp4_bits, "achieved": round(a_f4, 1), "peak": round(fp4_peak, 1), "frac": round(a_f4 / fp4_peak, 4),
                                    "frac_of_int8_equivalent_peak": round(a_f4 / int8_peak, 4), "unit": "TOPS",
                                    "limiter": "tcgen05.mma kind::mxf4 dispatch rate: 42-44 cycles per instruction for any N <= 64 measured alone, 50-57 under load (profiles/r2_mma_dispatch_rate_call20.txt, clock64 trace profiles/r2_tc05_trace_f4_n64_call26.txt): the MMA phases of consecutive items run back to back; ALU pipe 59-63 %, DRAM 44 %, tensor math 42-43 % active (ncu profiles/r2_tc05_ncu_summary_call31.txt)"}

    # ---- the reference's own measurement conventions, per bit width (both arms print these) -------------------------
    # profile_convention_tops: QGTC.bitMM2Bit_profile (QGTC_device.cu:379-434): 200 repetitions of ONE GEMM on the same
    # operands between two events (here one aliased batched launch of 200 problems -- L2-resident operands, as there);
    # single_gemm_per_launch_tops: 200 back-to-back launches of one GEMM each (the reference's literal protocol), captured
    # in a CUDA graph so that the host's ctypes overhead is not what is measured.
    prof_conv, single_launch = {}, {}
    for b in BITS:
        a0, x0 = A[0].contiguous(), X[b][0].contiguous()
        capi.bitMM2Bit_aliased(a0, x0, 200, M, K, N, 1, b, b, engine=engine)
        ms = event_timed(lambda: capi.bitMM2Bit_aliased(a0, x0, 200, M, K, N, 1, b, b, engine=engine), 3)
        prof_conv[str(b)] = round(OPS_PER_GEMM * 200 / (ms * 1e-3) / 1e12, 1)

        def many():
            for _ in range(200):
                capi.bitMM2Bit(a0, x0, M, K, N, 1, b, b, engine=engine)
        g1 = graph_of(many)
        ms = event_timed(g1.replay, 3)
        single_launch[str(b)] = round(OPS_PER_GEMM * 200 / (ms * 1e-3) / 1e12, 1)
        del g1

    # ---- end to end through the host-buffer C-ABI call (pinned host operands in, packed result read back): one pass over
    # all G graphs at the four bit widths; each group's adjacency is uploaded ONCE and used by every bit width;
    # uploads / GEMMs / downloads are pipelined by the library
    Ge = min(args.e2e_graphs, G)
    hA = A[:Ge].cpu().pin_memory()
    hX = [X[b][:Ge].cpu().pin_memory() for b in BITS]
    hO = [torch.empty((Ge, capi.packed_words(capi.LAYOUT_ROW, b, M, N)), dtype=torch.int32).pin_memory() for b in BITS]

    def e2e_step():
        capi.bitMM2Bit_host_batched_multi(hA, hX, Ge, M, K, N, 1, list(BITS), list(BITS), out_hosts=hO, engine=engine)

    e2e_step()
    barrier(world)
    e2e_steps = max(3, min(args.steps, 10))
    w0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier(world)
    e2e_rank_s = gather_floats((time.perf_counter() - w0) / e2e_steps, world)
    e2e_s = max(e2e_rank_s)
    e2e_val = OPS_PER_GEMM * Ge * len(BITS) * world / e2e_s / 1e12
    h2d = (hA.numel() + sum(x.numel() for x in hX)) * 4
    d2h = sum(o.numel() for o in hO) * 4
    e2e_ok = all(bool(torch.equal(hO[i][0], outs[b][0].cpu())) for i, b in enumerate(BITS))

    # spot check of the timed outputs against the oracle (not timed), graph 0, all four widths: the integer sums
    # (bitMM2Int: no saturation can hide a wrong sum) and the packed planes the timed launches wrote
    ok = None
    cpu = None
    if rank == 0:
        import numpy as np
        import c_oracle
        ok = True
        a_h = A[0].cpu().numpy()
        for b in BITS:
            x_h = X[b][0].cpu().numpy()
            gi = capi.bitMM2Int(A[0].contiguous(), X[b][0].contiguous(), M, K, N, 1, b, True, engine=engine).cpu().numpy()
            ok = ok and bool(np.array_equal(gi, c_oracle.bitMM2Int(a_h, x_h, M, K, N, 1, b, True)))
            want = c_oracle.bitMM2Bit(a_h, x_h, M, K, N, 1, b, b).reshape(-1)
            ok = ok and bool(np.array_equal(outs[b][0].cpu().numpy().view(np.uint32), want))
        if not args.no_cpu_baseline and world == 1:
            cpu = cpu_baseline_port(A[0].cpu().numpy(), {b: X[b][0].cpu().numpy() for b in BITS})

    epoch = None
    if not args.no_epoch:
        try:
            epoch = epoch_block(world, rank)
        except Exception as e:  # the headline metric must still be printed
            epoch = {"error": repr(e)[:300]}

    zero_tiles = None
    if rank == 0 and world == 1 and not args.no_zero_tile:
        try:
            del A, X, outs
            torch.cuda.empty_cache()
            zero_tiles = zero_tile_block()
        except Exception as e:
            zero_tiles = {"error": repr(e)[:300]}

    if rank == 0:
        cfg = config_dict(G)
        cfg.update({"engine": args.engine, "passes_per_step": P,
                    "timed_region": "K replays of one CUDA graph holding a step = %d passes x four batched calls (%d kernels)" % (P, 8 * P)})
        line = {
            "metric": "bit-GEMM int TOPS (1/2/4/8-bit) vs INT8 peak; Cluster-GCN epoch ms @1-8 GPU",
            "value": round(value, 3), "unit": "TOPS (2*M*K*N per GEMM)", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": "1-bit adjacency x b-bit features, exact integer results: b = 1, 2 -> E2M1 x E2M1 -> f32 (tcgen05 kind::mxf4, sums exact below "
                     "2^23, guarded); b = 4, 8 -> u8 x u8 -> s32 (tcgen05 kind::i8)",
            "data": "synthetic", "config": cfg,
            "rank_ms_timed_region": stats(rank_ms) if world > 1 else None,
            "per_bit_ms_per_call": {str(b): round(v, 4) for b, v in per_bit.items()},
            "per_bit_tops_per_call": {str(b): round(OPS_PER_GEMM * G / (v * 1e-3) / 1e12, 2) for b, v in per_bit.items()},
            "per_bit_ms_gemm_kernel": {str(b): round(v, 4) for b, v in per_bit_kernel.items()},
            "per_bit_tops_gemm_kernel": {str(b): round(v, 2) for b, v in tops_k.items()},
            "profile_convention_tops": prof_conv,
            "single_gemm_per_launch_tops": single_launch,
            "e2e": {"value": round(e2e_val, 3), "unit": "TOPS (2*M*K*N per GEMM)", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "qgtc_bitmm2bit_host_batched_multi (pinned host operands -> packed results on host; adjacency uploaded once per graph "
                           "for the four bit widths; upload/GEMM/download pipelined)",
                    "graphs_per_step": Ge * len(BITS), "steps": e2e_steps, "matches_device_result": e2e_ok, "numa_node_rank0": NUMA_NODE,
                    "rank_s_per_step": stats(e2e_rank_s) if world > 1 else None},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": round(achieved, 2), "peak": round(int8_peak, 1), "unit": "TFLOP/s",
                         "frac": round(achieved / int8_peak, 4), "traffic": traffic, "by_operand_path": by_path,
                         # the kernels are timed back to back inside a >= 100 ms region at the power cap (see clocks): against the
                         # SUSTAINED dense-GEMM figure of MEASURED_PEAKS.json (cuBLAS bf16 for 4 s, x 2 for INT8) the fraction is
                         "peak_sustained": round(2.0 * peaks["bf16_tflops_sustained"], 1),
                         "frac_of_sustained_peak": round(achieved / (2.0 * peaks["bf16_tflops_sustained"]), 4),
                         "note": "integer ops/s (2*M*K*N) of gemm_tc05_kernel alone (CUDA events around each of its launches, mean over the four bit "
                                 "widths; the call also runs prep_b_kernel: see per_bit_ms_per_call), against the INT8 peak the metric names = 2 x %s "
                                 "dense bf16 (burst, kernel timed alone; nominal 4500).  by_operand_path quotes each path against the peak of the "
                                 "instruction it issues (FP4 = 4 x bf16); the limiter fields say what paces each path (MMA dispatch rate on the 4-bit "
                                 "path, the three-stage TMEM ring on the u8 path)" % peaks["source"]},
            "cpu_baseline": cpu,
            "epochs": epoch,
            "zero_tile_sweep": zero_tiles,
            "clocks": clocks,
            "parity_spot_check": ok,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def zero_tile_block():
    """BASELINE configs[4] (Fig. 8b / 9): zero-tile jumping and adjacency-size sweep.  Synthetic adjacencies at partition sizes
    1K-16K -- i.i.d. density p and block-diagonal (128-node blocks, what a batch of small graphs looks like) -- times 64-wide 2-bit
    features, batches of G graphs.  Per case: dense-equivalent TOPS (2*M*K*N) without jumping, with on-the-fly jumping and with
    the precomputed tile bitmap; processed / total tiles as parse_counter.py:16 reports it (8 x 128-bit tiles of the reference's
    counters, and the 128 x 128 tiles the tensor core consumes); and a bit-exact check of graph 0 against the C oracle."""
    import numpy as np
    import torch
    import c_oracle
    from qgtc_ppopp22_b200 import capi
    Nz, b = 64, 2
    gen = torch.Generator(device="cuda").manual_seed(3)

    def timed(fn, reps=5):
        for _ in range(2):
            fn()
        return event_timed(fn, reps)

    rows = []
    for Mz in (1024, 4096, 16384):
        G = max(2, min(64, (256 << 20) // (Mz * Mz // 8)))
        X = torch.stack([capi.val2bit(torch.randint(0, 4, (Mz, Nz), device="cuda", generator=gen).float(), b, True, False).view(-1) for _ in range(G)])
        for kind, p in (("iid", 1.0), ("iid", 1e-2), ("iid", 1e-4), ("blockdiag128", 0.5)):
            adj = []
            for _ in range(G):
                if kind == "iid":
                    a = (torch.rand((Mz, Mz), device="cuda", generator=gen) < p).float()
                else:
                    a = torch.zeros((Mz, Mz), device="cuda")
                    blk = (torch.rand((Mz // 128, 128, 128), device="cuda", generator=gen) < p).float()
                    idx = torch.arange(Mz, device="cuda").view(Mz // 128, 128)
                    for k in range(Mz // 128):
                        a[idx[k][:, None], idx[k][None, :]] = blk[k]
                adj.append(capi.val2bit(a, 1, False, False).view(-1))
                del a
            A = torch.stack(adj)
            del adj
            total, nz = capi.tile_counts(A[0].view(capi.packed_shape(capi.LAYOUT_ROW, 1, Mz, Mz)), Mz, Mz, Nz, 1, b)
            bm = capi.tile_bitmap(A, Mz, Mz, G)
            out = torch.empty((G, capi.packed_words(capi.LAYOUT_ROW, b, Mz, Nz)), dtype=torch.int32, device="cuda")
            ops = 2.0 * Mz * Mz * Nz * G
            r = {"M": Mz, "K": Mz, "N": Nz, "bits": b, "pattern": kind, "density": p, "graphs": G,
                 "processed_over_total_tiles_8x128": round(nz / total, 4),
                 "processed_over_total_tiles_128x128": round(float(((bm.view(-1).cpu().numpy().view(np.uint32)[:, None] >> np.arange(32, dtype=np.uint32)) & 1).sum())
                                                             / (G * ((Mz + 127) // 128) ** 2), 4)}
            ms = timed(lambda: capi.bitMM2Bit_batched(A, X, G, Mz, Mz, Nz, 1, b, b, out=out, flags=capi.FLAG_NO_ZERO_JUMP))
            r["tops_no_jump"] = round(ops / (ms * 1e-3) / 1e12, 1)
            ms = timed(lambda: capi.bitMM2Bit_batched(A, X, G, Mz, Mz, Nz, 1, b, b, out=out))
            r["tops_jump_on_the_fly"] = round(ops / (ms * 1e-3) / 1e12, 1)
            ms = timed(lambda: capi.bitMM2Bit_tilemap(A, X, bm, G, Mz, Mz, Nz, b, b, out=out))
            r["tops_jump_tile_bitmap"] = round(ops / (ms * 1e-3) / 1e12, 1)
            want = c_oracle.bitMM2Bit(A[0].cpu().numpy(), X[0].cpu().numpy(), Mz, Mz, Nz, 1, b, b).reshape(-1)
            r["graph0_matches_oracle"] = bool(np.array_equal(out[0].cpu().numpy().view(np.uint32), want))
            rows.append(r)
            del A, bm, out
    return rows


def _epoch_args(drv, dataset, bits, extra, classes, hidden=16):
    base = ["--dataset", dataset, "--bit_width", str(bits), "--use_QGTC", "--n-classes", str(classes), "--n-hidden", str(hidden), "--n-epochs", "10"]
    return drv.build_parser().parse_args(base + extra)


def epoch_block(world, rank):
    """Second half of BASELINE.json's metric: forward-epoch ms of the Cluster-GCN / batched-GIN drivers on synthetic graphs of the
    reference's shapes, batches sharded across the ranks (strong scaling: the epoch is fixed, 75 or 32 batches).
    BASELINE configs[0]/[2]/[3]: ppi-shaped 2-bit; ogbn-arxiv-shaped at 1/2/4/8 bits; batched GIN 2-bit over 1024 small graphs.
    Modes: the op-major grouped mode (six launches per epoch, one CUDA graph; CUDA-event timing of a >= 100 ms region) is the
    product path; ppi also gets round 1's per-batch modes for continuity.  Rank 0 also times the torch restatement of the
    reference's DGL model on the host cores (reported baseline, not a target)."""
    from qgtc_ppopp22_b200 import cluster as C
    from qgtc_ppopp22_b200 import main_qgtc as drv
    out = {"n_gpus": world, "timing": "CUDA events around N graph replays (region >= 100 ms), max over ranks; per-batch eager mode: wall clock / epochs as main_qgtc.py:157-159"}
    G = ["--grouped", "--cuda-graph", "--min-region-ms", "100"]

    def rec(r):
        return {"epoch_ms": round(r["epoch_ms_device"], 5), "epochs_timed": r["epochs_timed"], "batches_this_rank": r["batches"],
                "rank_ms": stats(r["rank_epoch_ms_device"]) if world > 1 else None, "tile_bitmaps_used": r["tile_bitmaps_used"],
                "nonzero_tile_fraction": round(r["nonzero_tile_fraction"], 4) if r["nonzero_tile_fraction"] is not None else None}

    g = C.make_synthetic_graph("ppi", 1500)
    ppi = {"graph": "synthetic ppi-shaped: %d nodes, %d edges, %d feats, 121 classes; 75 batches of ~759 nodes, hidden 16, 2-bit" % (g.num_nodes, g.num_edges, g.feat.shape[1])}
    ppi["grouped_graph"] = rec(drv.run(_epoch_args(drv, "ppi", 2, G + (["--cpu-baseline"] if rank == 0 else []), 121), graph=g, quiet=True))
    r = drv.run(_epoch_args(drv, "ppi", 2, ["--grouped"], 121), graph=g, quiet=True)
    ppi["grouped_eager_ms"] = round(r["epoch_ms"], 4)
    r = drv.run(_epoch_args(drv, "ppi", 2, [], 121), graph=g, quiet=True)
    ppi["per_batch_eager_ms"] = round(r["epoch_ms"], 4)
    r = drv.run(_epoch_args(drv, "ppi", 2, ["--cuda-graph", "--streams", "16", "--min-region-ms", "100"], 121), graph=g, quiet=True)
    ppi["per_batch_graph_16_streams_ms"] = round(r["epoch_ms_device"], 5)
    ppi["reference_published_epoch_ms_sm86"] = 189.016   # README.md:88 (ppi, RTX 3090-class, for context only)
    out["cluster_gcn_ppi_2bit"] = ppi
    # the CPU baseline rides on the first ppi run (rank 0)
    a = _epoch_args(drv, "ppi", 2, G + ["--cpu-baseline"], 121)
    if rank == 0 and world == 1:
        rr = drv.run(a, graph=g, quiet=True)
        if "cpu_epoch_ms" in rr:
            out["cpu_baseline_epoch"] = {"epoch_ms": round(rr["cpu_epoch_ms"], 3), "cores": rr["cpu_cores"],
                                         "kind": "torch restatement of the DGL GraphSAGE forward (modules.py:9-45), fp32, CSR aggregation, ppi-shaped"}
    del g
    g = C.make_synthetic_graph("ogbn-arxiv", 1500)
    arx = {"graph": "synthetic ogbn-arxiv-shaped: %d nodes, %d edges, %d feats, 40 classes; 75 batches of ~2258 nodes, hidden 16" % (g.num_nodes, g.num_edges, g.feat.shape[1])}
    for b in (1, 2, 4, 8):
        arx["%d_bit" % b] = rec(drv.run(_epoch_args(drv, "ogbn-arxiv", b, G, 40), graph=g, quiet=True))
    arx["reference_published_epoch_ms_sm86"] = 208.616   # README.md:89
    out["cluster_gcn_arxiv"] = arx
    del g
    g = C.make_small_graph_batches()
    a = _epoch_args(drv, "small-graphs", 2, G + ["--run_GIN", "--psize", str(len(g.parts)), "--batch-size", "32"], 2, hidden=64)
    out["batched_gin_2bit"] = dict(rec(drv.run(a, graph=g, quiet=True)),
                                   graph="1024 synthetic small graphs (16-64 nodes), 29 feats, hidden 64; 32 batches of 32 graphs (block-diagonal, n ~ 1283)")
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
    """The reference arm.  The reference implements this path only in CUDA, so its own code path is the extension rebuilt for
    sm_100a from the unmodified sources (oracle/_ref/QGTC_ref), called through its public API (QGTC.bitMM2Bit on device tensors) on
    the same box: same workload definition as the repo's arm (batches of G graphs x four bit widths, one rank per GPU, MAX over
    ranks), one pass per step (a bounded sample: each of its calls costs milliseconds of host time, QGTC_device.cu:223-250).
    It also prints the reference's kernel-only figure in its own convention (bitMM2Bit_profile) and a same-box reference epoch
    (the logical-warp-index build, whose results are correct on a B200: oracle/ref_logical_warpid.h).  When the extension cannot
    be loaded or --ref-device cpu is given, the oracle's C port is timed on the host cores instead (rank 0 only)."""
    import numpy as np
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    G = args.graphs
    line_common = {
        "impl": "reference", "metric": "bit-GEMM int TOPS (1/2/4/8-bit) vs INT8 peak; Cluster-GCN epoch ms @1-8 GPU",
        "unit": "TOPS (2*M*K*N per GEMM)", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "data": "synthetic", "config": config_dict(G),
    }
    ref = None
    if args.ref_device in ("auto", "cuda"):
        try:
            import torch
            import build_ref
            if torch.cuda.is_available():
                ref = build_ref.load_variant("QGTC_ref")
        except Exception as e:
            ref = None
            if args.ref_device == "cuda":
                if rank == 0:
                    print(json.dumps({"impl": "reference", "unavailable": "reference extension not loadable: %r" % (e,)}))
                return
    if ref is not None:
        import torch
        rank, local, world = dist_setup(args.gpus)
        gen = torch.Generator(device="cuda").manual_seed(3 + rank)
        A, X = [], {b: [] for b in BITS}
        for g in range(G):
            A.append(ref.val2bit((torch.rand((M, K), device="cuda", generator=gen) < 0.5).float(), 1, False, False))
            for b in BITS:
                X[b].append(ref.val2bit(torch.randint(0, 1 << b, (K, N), device="cuda", generator=gen).float(), b, True, False))

        def step():
            for b in BITS:
                for g in range(G):
                    ref.bitMM2Bit(A[g], X[b][g], M, K, N, 1, b, b)

        for _ in range(max(1, min(args.warmup, 2))):
            step()
        barrier(world)
        w0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        torch.cuda.synchronize()
        barrier(world)
        rank_s = gather_floats((time.perf_counter() - w0) / args.steps, world)
        s = max(rank_s)
        # kernel-only figures of the reference in its own convention: bitMM2Bit_profile (200 back-to-back launches, prints TFLOPs)
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
        del A, X
        torch.cuda.empty_cache()
        # same-box reference epoch: the reference's ops driven by this repo's mirror of main_qgtc.py (per-batch loop, eager),
        # ppi-shaped synthetic graph, batches sharded over the ranks like the repo's arm
        epoch = None
        if not args.no_epoch:
            try:
                import build_ref as _br
                from qgtc_ppopp22_b200 import cluster as C
                from qgtc_ppopp22_b200 import main_qgtc as drv
                ref_lw = _br.load_variant("QGTC_ref_lw")
                g = C.make_synthetic_graph("ppi", 1500)
                a = drv.build_parser().parse_args(["--dataset", "ppi", "--n-epochs", "3", "--bit_width", "2", "--use_QGTC", "--n-classes", "121"])
                r = drv.run(a, graph=g, quiet=True, ops=ref_lw)
                epoch = {"cluster_gcn_ppi_2bit": {"per_batch_eager_ms": round(r["epoch_ms"], 3), "epochs_timed": r["epochs_timed"],
                                                  "ops": "oracle/_ref/QGTC_ref_lw (reference sources, logical warp index; results correct on B200)",
                                                  "n_gpus": world}}
            except Exception as e:
                epoch = {"error": repr(e)[:300]}
        if rank == 0:
            val = OPS_PER_GEMM * G * len(BITS) * world / s / 1e12
            line = dict(line_common)
            line.update({"value": round(val, 3), "ms_per_step": round(s * 1e3, 3),
                         "dtype": "b1 AND+popc (wmma m8n8k128; emulated with IMMA on sm_100a)",
                         "step_definition": "one pass over the batch of %d graphs at each of the four bit widths through QGTC.bitMM2Bit of the "
                                            "unmodified reference rebuilt for sm_100a, one process per GPU" % G,
                         "rank_s_per_step": stats(rank_s) if world > 1 else None,
                         "cpu_baseline": {"value": round(val, 3), "unit": "TOPS (2*M*K*N per GEMM)", "cores": 0, "kind": "reference",
                                          "sample": "reference CUDA kernels on the same B200s, one pass (%d graphs x 4 bit widths) per step "
                                                    "(the reference has no CPU implementation of this path)" % G},
                         "profile_convention_tops": kernel_only,
                         "single_gemm_per_launch_tops": kernel_only,
                         "reference_kernel_only_tops": kernel_only,
                         "epochs": epoch,
                         "e2e": {"value": round(val, 3), "unit": "TOPS (2*M*K*N per GEMM)", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            print(json.dumps(line))
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    if rank != 0:
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
    line.update({"value": round(val, 4), "ms_per_step": round(s * 1e3, 3), "dtype": "b1 AND+popcount (uint64)", "n_gpus": 1,
                 "step_definition": "one graph x {1,2,4,8} bits per step, oracle C port on the host cores",
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
