#!/usr/bin/env python
"""bench.py -- headline benchmark: atoms/sec, ResidueLevel, 100 sphere points, probe 1.4 A, ProtOr radii,
on the synthetic AlphaFold-like proteome batch (BASELINE.json configs[1]: 4,400 structures of ~2.4k atoms).

    python bench.py --gpus N --steps K --warmup W            # GPU arm (torchrun for N > 1, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle's C restatement of RustSASA's
                                                             # CPU path on the host cores (rank 0 only)

One "step" = one pass of the hot path over one batch.  `value` = atoms/s with the batch resident in HBM
(CUDA events on the launching stream, max over ranks); `e2e` = the same metric through the host-buffer C-ABI
call (pinned host buffers in and out, H2D + kernels + D2H inside the timed region).  Multi-GPU is weak
scaling: every rank processes its own seeded batch of the same shape; there is no data-path collective.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "atoms_per_sec_residue_level_100pts"
UNIT = "atoms/s"
N_POINTS = 100
PROBE = 1.4
FLOP_PER_TEST = 6.0       # 1 FMUL + 2 FFMA + 1 FSETP per point-neighbour test
FLOP_PER_PAIR = 20.0      # per-pair setup (3 sub, vmag, limit incl. the division)
# algorithmic bytes per atom of THIS run (SURVEY.md 8d lists 16 B float4 in + 4 B count + 4 B SASA + ~0.5 B residue sum; a
# ResidueLevel run writes neither the counts nor the per-atom areas): 16 B in + 4 B per residue sum out
BYTES_IN_PER_ATOM = 16.0


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def peaks():
    p = dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source="fallback (B200_PROFILING.md)")
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            m = json.load(fh)
        p = dict(hbm_gbs=float(m["hbm_gbs"]), sm_max_mhz=float(m.get("sm_max_mhz", 1965.0)), source="MEASURED_PEAKS.json")
    except Exception:
        pass
    return p


def ncu_traffic(n_structures):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this same command
    (profiles/ncu_traffic.json, written by tools/ncu_summary.py --traffic); None when the workload differs."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as fh:
            t = json.load(fh)
        if int(t.get("structures", -1)) != int(n_structures):
            return None, None
        return float(t["dram_bytes_per_step"]), t
    except Exception:
        return None, None


def host_threads():
    """Every host core this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to its
    workers, which would time the CPU arm on one thread at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(data, seconds_target=15.0, threads=0):
    """Time the oracle's -O3 build (C restatement of RustSASA's CPU path, directory-mode threading: one structure
    per task over all host cores) on a bounded prefix of the batch.  Returns (atoms/s, info)."""
    from oracle import load
    fast = load(fast=True)
    cores = threads or host_threads()
    S = data.n_structures
    probe_n = min(S, max(2 * cores, 16))
    a1 = int(data.struct_off[probe_n])
    t0 = time.perf_counter()
    fast.run_batch(data.xyzr[:a1], data.struct_off[:probe_n + 1], PROBE, N_POINTS, 8, cores, want_counts=False)
    rate = a1 / (time.perf_counter() - t0)
    want_atoms = rate * seconds_target
    n = int(np.searchsorted(data.struct_off, want_atoms, side="left"))
    n = int(min(S, max(probe_n, n)))
    a1 = int(data.struct_off[n])
    g1 = int(data.struct_seg_off[n])
    t0 = time.perf_counter()
    out = fast.run_batch(data.xyzr[:a1], data.struct_off[:n + 1], PROBE, N_POINTS, 8, cores,
                         seg_be=data.seg_be[:g1], struct_seg_off=data.struct_seg_off[:n + 1], want_counts=True)
    dt = time.perf_counter() - t0
    # k-bar of the reference's neighbour definition on a small sub-sample (for the roofline's algorithmic flops)
    m = min(n, 48)
    ks, na = 0.0, 0
    for s in range(m):
        b0, b1 = int(data.struct_off[s]), int(data.struct_off[s + 1])
        r = fast.calculate_sasa_internal(data.xyzr[b0:b1], PROBE, N_POINTS, want_k=True)
        ks += float(r["k"].sum())
        na += b1 - b0
    info = dict(value=a1 / dt, unit=UNIT, cores=cores, kind="port",
                sample=f"first {n} of {S} structures ({a1} atoms), {dt:.1f} s, one structure per task on {cores} threads "
                       f"(C restatement of RustSASA's CPU path, -O3 -march=native; the Rust crate cannot be built here)",
                k_mean=ks / max(na, 1), exposed_fraction=float(out["counts"].sum()) / (N_POINTS * a1))
    return info, out, (n, a1, g1)


def run_reference(args, rank):
    """CPU arm: rank 0 alone works; the other ranks exit 0.  Every step is the WHOLE batch of the GPU arm's configuration
    (3 s on 16 cores), the oracle's -O3 build with the reference's directory-mode threading on all host cores."""
    if rank != 0:
        return
    from rustsasa_b200 import workloads as W
    data = W.proteome_batch(args.structures)
    from oracle import load
    fast = load(fast=True)
    cores = host_threads()
    S, N, G = data.n_structures, data.n_atoms, int(data.seg_be.shape[0])
    times = []
    out = None
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        out = fast.run_batch(data.xyzr, data.struct_off, PROBE, N_POINTS, 8, cores, seg_be=data.seg_be,
                             struct_seg_off=data.struct_seg_off, want_counts=(i == 0))
        if i >= args.warmup:
            times.append(time.perf_counter() - t0)
    dt = float(np.mean(times))
    v = N / dt
    info = dict(value=v, unit=UNIT, cores=cores, kind="port",
                sample=f"all {S} structures ({N} atoms) per step, {dt:.1f} s, one structure per task on {cores} threads "
                       f"(C restatement of RustSASA's CPU path, -O3 -march=native; the Rust crate cannot be built here)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(args, data, args.gpus),
        "cpu_baseline": info,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def bench_config(args, data, world):
    """The same dict in both arms (the driver compares them)."""
    return {"workload": workload_name(args), "level": "residue", "n_points": N_POINTS, "probe": PROBE, "radii": "ProtOr",
            "structures": data.n_structures, "atoms": data.n_atoms, "residues": int(data.seg_be.shape[0]),
            "parallelism": f"structures sharded over {world} GPU(s), no collective; every rank its own seeded batch of this shape",
            "l2": "inputs (16 B x atoms = %.0f MB per step) exceed the 126 MB L2; no flush needed" % (data.n_atoms * 16 / 1e6)}


def workload_name(args):
    return (f"cfg2 synthetic AlphaFold-like proteome batch: {args.structures} structures of ~2.4k atoms "
            f"(fragments of the reference's tests/data coordinate sets, rigid transform + 0.05 A jitter), seed 20261017")


def bind_rank_to_cores(local_rank, world):
    """N > 1: give every rank its own slice of the host cores near its GPU (NVML's ideal-CPU mask), before any pinned buffer is
    allocated, so that eight host pipelines neither migrate nor share cores (e2e efficiency 0.875 at 8 GPUs with
    unbound ranks).  Returns a short description for the JSON line."""
    if world <= 1:
        return None
    try:
        allowed = sorted(os.sched_getaffinity(0))
        near = allowed
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
            words = pynvml.nvmlDeviceGetCpuAffinity(h, (max(allowed) // 64) + 1)
            mask = [64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1]
            near = [c for c in allowed if c in set(mask)] or allowed
        except Exception:
            pass
        per = max(1, len(near) // world)
        mine = near[(local_rank * per) % len(near):][:per] or near
        os.sched_setaffinity(0, mine)
        return f"rank bound to {len(mine)} cores {mine[0]}-{mine[-1]}"
    except Exception as e:   # affinity is an optimisation, never a failure
        return f"unbound ({e})"


def sha_counts(counts):
    import hashlib
    return hashlib.sha256(np.ascontiguousarray(counts, dtype="<u4").tobytes()).hexdigest()


def secondary_configs(args, eng, rank, world, local_rank, weak_seg_rank0):
    """BASELINE.json configs 2 (strong scaling), 3, 4 and 5 measured after the headline legs; every entry carries its own
    parity flag.  Times are CUDA events / host clocks bracketed by barriers, max over ranks.  Returned on rank 0."""
    import torch
    import torch.distributed as dist
    from rustsasa_b200 import workloads as W
    from rustsasa_b200.engine import BatchResult
    from rustsasa_b200.shard import partition_structures, run_atom_range, take_shard
    out = {}
    reps = max(3, min(args.steps, 10))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allr(flag):
        t = torch.tensor([1.0 if flag else 0.0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t[0] > 0.5)

    def timed_events(fn, n):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        return maxr(e0.elapsed_time(e1) / n)

    def timed_wall(fn, n):
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / n
        barrier()
        return maxr(dt * 1e3)

    with open(os.path.join(ROOT, "tests", "golden", "cfg_hashes.json")) as fh:
        golden = json.load(fh)

    # ---- cfg2, STRONG scaling: the one seeded batch of the N = 1 run cut into contiguous cost-balanced shards -------------
    try:
        d = W.proteome_batch(args.structures, seed=W.SEED)
        bounds = partition_structures(d.struct_off, world, N_POINTS)
        sh = take_shard(bounds, rank, d.xyzr, d.struct_off, d.seg_be, d.struct_seg_off, d.seg_polar)
        b = eng.batch(sh.struct_off, sh.seg_be, sh.struct_seg_off, sh.seg_polar)
        na, ng = sh.a1 - sh.a0, sh.g1 - sh.g0
        d_x = torch.from_numpy(np.ascontiguousarray(sh.xyzr)).cuda()
        d_seg = torch.zeros(max(ng, 1), dtype=torch.float32, device="cuda")
        h_x = eng.pinned_empty((na, 4), np.float32)
        h_x[...] = sh.xyzr
        res = BatchResult(seg_sasa=eng.pinned_empty(max(ng, 1), np.float32))
        run_d = lambda: b.run_device(d_x, seg_sasa=d_seg, probe_radius=PROBE, n_points=N_POINTS)   # noqa: E731
        run_h = lambda: b.run_host(h_x, probe_radius=PROBE, n_points=N_POINTS, result=res)          # noqa: E731
        for _ in range(2):
            run_d()
            run_h()
        dev_ms = timed_events(run_d, reps)
        e2e_ms = timed_wall(run_h, reps)
        # results gathered on rank 0 (they sit in host memory; sizes are known from the partition)
        G = int(d.seg_be.shape[0])
        gathered = None
        t_g = time.perf_counter()
        if world > 1:
            sizes = [int(d.struct_seg_off[bounds[r + 1]] - d.struct_seg_off[bounds[r]]) for r in range(world)]
            pad = max(sizes)
            mine = torch.zeros(pad, dtype=torch.float32, device="cuda")
            mine[:ng] = torch.from_numpy(np.asarray(res.seg_sasa)[:ng].copy()).cuda()
            allb = torch.empty(pad * world, dtype=torch.float32, device="cuda")
            dist.all_gather_into_tensor(allb, mine)
            if rank == 0:
                a = allb.cpu().numpy()
                gathered = np.concatenate([a[r * pad: r * pad + sizes[r]] for r in range(world)])
        else:
            gathered = np.asarray(res.seg_sasa)[:ng].copy()
        gather_ms = (time.perf_counter() - t_g) * 1e3
        ok = True
        if rank == 0:
            ok = gathered.shape[0] == G and (weak_seg_rank0 is None or bool(np.array_equal(gathered, weak_seg_rank0)))
        out["cfg2_strong"] = {
            "workload": "the seeded 4,400-structure batch of the N = 1 run, cut by shard.partition_structures into contiguous "
                        "cost-balanced shards, one per GPU; no data-path collective, residue sums gathered on rank 0",
            "scaling": "strong", "structures": d.n_structures, "atoms": d.n_atoms, "value": d.n_atoms / (dev_ms * 1e-3),
            "ms_per_step": dev_ms, "e2e": {"value": d.n_atoms / (e2e_ms * 1e-3), "ms_per_step": e2e_ms},
            "gather_ms": gather_ms, "unit": UNIT,
            "parity": allr(ok), "parity_against": "bit-identical to the single-GPU residue sums of the same batch",
            "limiter": "per-GPU work shrinks to 1/N of a 6.4 ms step while launch ramp-up, the largest-first tail of one launch and "
                       "the host call's fixed cost (~0.15 ms) stay"}
        b.close()
        del d_x, d_seg, h_x, res, d
    except Exception as e:   # never lose the headline line to a secondary measurement
        out["cfg2_strong"] = {"error": repr(e)}

    # ---- cfg3: MD trajectory, ProteinLevel, frames sharded over the ranks ---------------------------------------------------
    try:
        F_all = 10000
        md = W.md_trajectory(n_frames=F_all, n_atoms=5000)
        NA = md.xyz.shape[1]
        f0, f1 = F_all * rank // world, F_all * (rank + 1) // world
        F = f1 - f0
        G = len(md.seg_be)
        off = np.arange(F + 1, dtype=np.uint64) * NA
        b = eng.batch(off, np.tile(md.seg_be, (F, 1)), np.arange(F + 1, dtype=np.uint64) * G, np.tile(md.seg_polar, F))
        h_xyz = eng.pinned_empty((F * NA, 3), np.float32)
        h_xyz[...] = md.xyz[f0:f1].reshape(-1, 3)
        res = BatchResult(protein=eng.pinned_empty((F, 3), np.float32))
        run_h = lambda: b.run_frames_host(h_xyz, md.radii, want=("protein",), result=res)   # noqa: E731
        run_h()
        e2e_ms = timed_wall(run_h, 3)
        d_xyzr = torch.from_numpy(np.concatenate([md.xyz[f0:f1].reshape(-1, 3), np.tile(md.radii, F)[:, None]], axis=1).astype(np.float32)).cuda()
        d_prot = torch.zeros((F, 3), dtype=torch.float32, device="cuda")
        run_d = lambda: b.run_device(d_xyzr, protein=d_prot)   # noqa: E731
        run_d()
        dev_ms = timed_events(run_d, 3)
        # parity: the oracle's protein totals of this rank's first frame (tests/golden/cfg_hashes.json, written by
        # tools/make_cfg_hashes.py for the first frames of every rank at N = 1, 2, 4, 8), bit-identical; host leg == device leg
        g3 = golden["cfg3"]
        ok = g3["frames"] == F_all and g3["atoms_per_frame"] == NA
        want = g3["protein_totals"].get(str(f0))
        if want is not None:
            ok = ok and bool(np.array_equal(np.asarray(res.protein)[0], np.asarray(want, np.float32)))
        ok = ok and bool(np.array_equal(np.asarray(res.protein), d_prot.cpu().numpy()))
        m = 1 if want is not None else 0
        out["cfg3_md_frames"] = {
            "workload": f"{F_all} frames x {NA} atoms, ProteinLevel, 100 points, frames sharded over {world} GPU(s); 12 B/atom/frame "
                        "on the wire, radii sent once", "scaling": "strong", "atoms": F_all * NA,
            "value": F_all * NA / (dev_ms * 1e-3), "ms_per_step": dev_ms,
            "e2e": {"value": F_all * NA / (e2e_ms * 1e-3), "ms_per_step": e2e_ms, "frames_per_s": F_all / (e2e_ms * 1e-3),
                    "h2d_bytes_per_step": F_all * NA * 12, "d2h_bytes_per_step": F_all * 12},
            "unit": UNIT, "parity": allr(ok),
            "parity_against": f"oracle protein totals of the first frame of every rank (stored fingerprints; {m} checked here), "
                              "bit-identical; host leg == device leg on all frames"}
        b.close()
        del d_xyzr, d_prot, h_xyz, res, md
    except Exception as e:
        out["cfg3_md_frames"] = {"error": repr(e)}

    # ---- cfg4: one 150k-atom assembly, AtomLevel, 100 points, single GPU (every rank runs it; rank 0 reports) ------------
    try:
        a = W.large_assembly(150000)
        b = eng.batch(a.struct_off)
        NA = a.n_atoms
        h = eng.pinned_empty((NA, 4), np.float32)
        h[...] = a.xyzr
        res = BatchResult(atom_sasa=eng.pinned_empty(NA, np.float32), counts=eng.pinned_empty(NA, np.uint32))
        run_h = lambda: b.run_host(h, n_points=100, want=("counts", "atom"), result=res)   # noqa: E731
        d_x = torch.from_numpy(a.xyzr).cuda()
        d_atom = torch.zeros(NA, dtype=torch.float32, device="cuda")
        run_d = lambda: b.run_device(d_x, n_points=100, atom_sasa=d_atom)   # noqa: E731
        for _ in range(3):
            run_h()
            run_d()
        e2e_ms = timed_wall(run_h, 10)
        dev_ms = timed_events(run_d, 10)
        launches = b.sync()["gpu_launches"]
        g = golden["cfg4"]
        ok = NA == g["atoms"] and sha_counts(np.asarray(res.counts)) == g["sha256_counts"]
        out["cfg4_assembly"] = {"workload": f"one {NA}-atom globule, AtomLevel, 100 points, one GPU", "atoms": NA,
                                "value": NA / (dev_ms * 1e-3), "ms_per_step": dev_ms, "launches_per_step": int(launches),
                                "e2e": {"value": NA / (e2e_ms * 1e-3), "ms_per_step": e2e_ms}, "unit": UNIT, "parity": allr(ok),
                                "parity_against": "sha256 of the oracle's per-atom counts (tests/golden/cfg_hashes.json), exact"}
        b.close()
        del d_x, d_atom, h, res, a
    except Exception as e:
        out["cfg4_assembly"] = {"error": repr(e)}

    # ---- cfg5: 1M-atom capsid, 960 points, atom ranges split over the ranks + one all-reduce ---------------------------------
    try:
        a = W.capsid_shell(1000000)
        b = eng.batch(a.struct_off)
        NA = a.n_atoms
        d_x = torch.from_numpy(a.xyzr).cuda()
        counts = torch.empty(NA, dtype=torch.int32, device="cuda")
        atom = torch.empty(NA, dtype=torch.float32, device="cuda")

        def kernels_only():
            b.run_atom_range_device(d_x, rank, world, n_points=960, counts=counts, atom_sasa=atom)

        def step():
            run_atom_range(lambda r, w: (kernels_only(), (counts, atom))[1], rank, world)
        for _ in range(3):
            step()
        k_ms = timed_events(kernels_only, 5)
        dev_ms = timed_events(step, 5)
        wall_ms = timed_wall(step, 5)
        step()
        torch.cuda.synchronize()
        g = golden["cfg5"]
        got = counts.cpu().numpy().view(np.uint32)
        ok = NA == g["atoms"] and int(got.astype(np.int64).sum()) == g["sum_counts"] and sha_counts(got) == g["sha256_counts"]
        # the exchange fused into the kernel: every rank writes its atoms' values into all ranks' vectors over NVLink
        # (torch symmetric memory), no zero-fill, no all-reduce; a barrier on either side
        fused = None
        if world > 1:
            try:
                from rustsasa_b200.shard import PeerVectors, run_atom_range_peers
                pv = PeerVectors(NA)

                def step_peers():
                    run_atom_range_peers(lambda r, w, v: b.run_atom_range_peers_device(d_x, r, w, v.count_ptrs, v.atom_ptrs, n_points=960), pv)
                for _ in range(3):
                    step_peers()
                p_ms = timed_events(step_peers, 5)
                p_wall = timed_wall(step_peers, 5)
                pv.counts.zero_()
                torch.cuda.synchronize()
                barrier()
                step_peers()
                torch.cuda.synchronize()
                got_p = pv.counts.cpu().numpy().view(np.uint32)
                okp = bool(np.array_equal(got_p, got))
                fused = {"value": NA / (p_ms * 1e-3), "ms_per_step": p_ms, "wall_ms_per_step": p_wall, "parity": allr(okp),
                         "note": "sasa_b200_batch_run_atom_range_peers_device: peer stores from inside the atoms kernel + two barriers"}
            except Exception as e:
                fused = {"error": repr(e)}
        out["cfg5_capsid"] = {
            "workload": f"one {NA}-atom capsid shell, AtomLevel, 960 points; every rank builds the whole cell list and evaluates its "
                        f"interleaved share of the cell-sorted atoms; per-rank count / area vectors summed by ncclAllReduce ({world} GPU(s))",
            "scaling": "strong", "atoms": NA, "value": NA / (dev_ms * 1e-3), "ms_per_step": dev_ms,
            "kernels_only": {"value": NA / (k_ms * 1e-3), "ms_per_step": k_ms},
            "wall": {"value": NA / (wall_ms * 1e-3), "ms_per_step": wall_ms}, "unit": UNIT, "parity": allr(ok),
            "parity_against": "sha256 + sum of the oracle's per-atom counts at full size (tests/golden/cfg_hashes.json), exact",
            "sum_counts": int(got.astype(np.int64).sum())}
        if fused is not None:
            out["cfg5_capsid"]["peer_writes"] = fused
        b.close()
    except Exception as e:
        out["cfg5_capsid"] = {"error": repr(e)}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--structures", type=int, default=4400)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-secondary", action="store_true", help="skip the cfg2-strong / cfg3 / cfg4 / cfg5 block")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    world = env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    binding = bind_rank_to_cores(local_rank, world)
    import torch
    import torch.distributed as dist
    from rustsasa_b200 import Engine
    from rustsasa_b200 import workloads as W

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    data = W.proteome_batch(args.structures, seed=W.SEED + rank)     # weak scaling: one full batch per rank
    eng = Engine(local_rank)
    batch = eng.batch(data.struct_off, data.seg_be, data.struct_seg_off, data.seg_polar)
    N, G = data.n_atoms, int(data.seg_be.shape[0])

    # ---- device-resident leg (value) ---------------------------------------------------------------
    d_xyzr = torch.from_numpy(data.xyzr).cuda()
    d_seg = torch.zeros(G, dtype=torch.float32, device="cuda")
    # the kernels are launched on this (non-default) torch stream, and so are the timing events
    tstream = torch.cuda.Stream()
    torch.cuda.synchronize()
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device():
        batch.run_device(d_xyzr, seg_sasa=d_seg, probe_radius=PROBE, n_points=N_POINTS, stream=stream)

    for _ in range(args.warmup):
        step_device()
    barrier()
    launches_per_step = batch.sync()["gpu_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    ev[0].record()
    for i in range(args.steps):
        step_device()
        ev[i + 1].record()
    barrier()
    dev_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    stats_dev = batch.sync()

    # ---- end-to-end leg (e2e): pinned host buffers through the host C-ABI call ------------------------
    h_xyzr = eng.pinned_empty((N, 4), np.float32)
    h_xyzr[...] = data.xyzr
    from rustsasa_b200.engine import BatchResult
    h_res = BatchResult(seg_sasa=eng.pinned_empty(G, np.float32))
    for _ in range(args.warmup):
        batch.run_host(h_xyzr, probe_radius=PROBE, n_points=N_POINTS, result=h_res)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        batch.run_host(h_xyzr, probe_radius=PROBE, n_points=N_POINTS, result=h_res)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    clocks = sampler.stop()
    stats_e2e = h_res.stats

    # ---- the same with the indexed-radius wire format (12 B coordinates + 1 B palette index per atom) ---------------------
    from rustsasa_b200.engine import index_radii
    e2e_idx_s, same_idx, pal = None, None, index_radii(data.xyzr[:, 3])
    if pal is not None:
        h_xyz3 = eng.pinned_empty((N, 3), np.float32)
        h_xyz3[...] = data.xyzr[:, :3]
        h_idx = eng.pinned_empty(N, np.uint8)
        h_idx[...] = pal[1]
        h_res2 = BatchResult(seg_sasa=eng.pinned_empty(G, np.float32))
        for _ in range(args.warmup):
            batch.run_indexed_host(h_xyz3, h_idx, pal[0], probe_radius=PROBE, n_points=N_POINTS, result=h_res2)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            batch.run_indexed_host(h_xyz3, h_idx, pal[0], probe_radius=PROBE, n_points=N_POINTS, result=h_res2)
        torch.cuda.synchronize()
        e2e_idx_s = time.perf_counter() - t0
        barrier()
        same_idx = bool(np.array_equal(np.asarray(h_res2.seg_sasa), np.asarray(h_res.seg_sasa)))
        del h_xyz3, h_idx

    # results of the two legs must agree bit for bit
    same = bool(np.array_equal(d_seg.cpu().numpy(), np.asarray(h_res.seg_sasa)))

    # ---- max over ranks --------------------------------------------------------------------------------
    t = torch.tensor([dev_ms, e2e_s * 1e3, (e2e_idx_s or 0.0) * 1e3], dtype=torch.float64, device="cuda")
    per_rank = [t.clone() for _ in range(world)] if world > 1 else [t]
    if world > 1:
        dist.all_gather(per_rank, t)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms_max, e2e_ms_max, e2e_idx_ms_max = float(t[0]), float(t[1]), float(t[2])
    tot = torch.tensor([float(N)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    atoms_all = float(tot[0])

    weak_seg = np.asarray(h_res.seg_sasa).copy() if rank == 0 else None
    torch.cuda.set_stream(torch.cuda.default_stream())
    secondary = None
    if not args.no_secondary:
        # free the headline buffers first (the secondary configurations bring their own)
        del d_xyzr, d_seg
        secondary = secondary_configs(args, eng, rank, world, local_rank, weak_seg)

    if rank == 0:
        value = atoms_all * args.steps / (dev_ms_max * 1e-3)
        e2e = atoms_all * args.steps / (e2e_ms_max * 1e-3)
        pk = peaks()
        props = torch.cuda.get_device_properties(local_rank)
        peak_tflops = props.multi_processor_count * 128 * 2 * pk["sm_max_mhz"] * 1e6 / 1e12
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(args, data, world),
            "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": int(N * 16), "d2h_bytes_per_step": int(G * 4),
                    "ms_per_step": e2e_ms_max / args.steps, "timing": "host wall clock around the synchronous C-ABI call",
                    "wire_format": "float4 {x, y, z, r} per atom (sasa_b200_batch_run_host)"},
            "gpu_launches": int(launches_per_step) * args.steps,
            "clocks": clocks,
            "results_identical_device_vs_host_leg": same,
            "step_ms": [round(x, 4) for x in step_ms],
            "gpu_stats": {"tight_neighbours_per_atom": stats_dev["neighbor_pairs"] / max(1, N * args.steps),
                          "streamed_atoms": stats_dev["streamed_atoms"], "launches_per_step": int(launches_per_step),
                          "e2e_device_span_ms": stats_e2e["kernel_ms"]},
        }
        if e2e_idx_s is not None:
            # The end-to-end figure is the indexed-radius call: it is what the extraction step emits for ProtOr radii (ten
            # distinct values), its results are bit-identical, and with several GPUs pulling over PCIe at once the host copy is
            # the limiter.  The float4 call of the same batch stays next to it.
            out["e2e_float4"] = out["e2e"]
            out["e2e"] = {
                "value": atoms_all * args.steps / (e2e_idx_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(N * 13),
                "d2h_bytes_per_step": int(G * 4), "ms_per_step": e2e_idx_ms_max / args.steps,
                "timing": "host wall clock around the synchronous C-ABI call", "palette": int(pal[0].shape[0]),
                "results_identical_to_float4_form": same_idx,
                "wire_format": "12 B coordinates + 1 B radius-palette index per atom (sasa_b200_batch_run_indexed_host)"}
        if world > 1:
            out["per_rank_ms_per_step"] = {"device": [round(float(x[0]) / args.steps, 3) for x in per_rank],
                                           "e2e": [round(float(x[1]) / args.steps, 3) for x in per_rank],
                                           "e2e_indexed": [round(float(x[2]) / args.steps, 3) for x in per_rank]}
        if binding:
            out["host_binding"] = binding
        cpu = None
        if not args.no_cpu and world == 1:   # the CPU leg is an N = 1 measurement (rank 0 alone would stall the other ranks)
            cpu, cpu_out, (n, a1, g1) = cpu_sample(data)
            out["cpu_baseline"] = cpu
            out["parity_on_cpu_sample"] = bool(np.array_equal(cpu_out["seg"], weak_seg[:g1]))
        # without the CPU leg: the oracle's figure for this same seeded batch from the committed N = 1 run (profiles/r01h_bench.json)
        k_mean = cpu["k_mean"] if cpu else 43.354429297893255
        flops_per_atom = FLOP_PER_TEST * N_POINTS * k_mean + FLOP_PER_PAIR * k_mean
        per_gpu_step_s = dev_ms_max * 1e-3 / args.steps
        traffic, traffic_info = ncu_traffic(args.structures)
        achieved = N * flops_per_atom / per_gpu_step_s / 1e12
        alg_bytes = N * BYTES_IN_PER_ATOM + G * 4.0    # what THIS run reads and writes: float4 atoms in, residue sums out
        out["roofline"] = {
            "bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
            "traffic": traffic,
            "traffic_unit": "DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum over the step's launches)",
            "traffic_source": (traffic_info or {}).get("source"),
            "algorithmic_bytes_per_step": alg_bytes,
            "kernel": "sasa_tight_kernel (fused per-structure kernel; all launches of a step)",
            "note": ("FP32 CUDA-core compare roofline (SURVEY.md 8d): algorithmic flops = atoms x (6 x n_points + 20) x k_mean, "
                     f"k_mean = {k_mean:.2f} reference-definition neighbours/atom measured by the oracle on a sample; "
                     f"peak = SMs x 128 x 2 x {pk['sm_max_mhz']:.0f} MHz ({pk['source']}); early exit may legitimately push frac past "
                     "what the executed-instruction count implies"),
            "hbm": {"achieved": alg_bytes / per_gpu_step_s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": alg_bytes / per_gpu_step_s / 1e9 / pk["hbm_gbs"], "bytes_per_atom": alg_bytes / N,
                    "note": "16 B float4 per atom in + 4 B per residue sum out (a ResidueLevel run writes no per-atom output)"},
        }
        if secondary is not None:
            out["secondary"] = secondary
        print(json.dumps(out))
    batch.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
