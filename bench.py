#!/usr/bin/env python
"""bench.py -- the `blamm scan` hot path on N B200s (one process per GPU, launched by torchrun for N > 1).

--config c2 (default; BASELINE.json configs[1]): JASPAR-LIKE set of 900 PWMs (real JASPAR CORE is not available offline;
  seeded Dirichlet-multinomial count matrices with CORE-like lengths, blamm_b200/synth.py) x both strands = 1800 columns,
  `-rc -pt 1e-4` thresholds from theoretical histograms, against ONE synthetic stream of N x 100 Mbp dealt in 100 Mbp chunks
  (+ maxLen-1 halo) to the N ranks (blamm_b200/shard.py: chunk k -> rank k mod N; weak scaling, no collective on the data
  path; the ranks exchange only hit counts for the stream-order merge).  A step = one pass of the hot path over the rank's chunk.
    value : device-resident input (2-bit codes already in HBM), CUDA events around the scoring kernels (tensor-core filter,
            expand, exact rescore, device-side (position, column) ordering), L2 flushed between steps, max over ranks.
    e2e   : the same chunk as CHARACTERS in pinned host memory through the product's hand-over: host 2-bit packer
            (blamm_pack_ascii on the rank's share of the host cores) -> b200scan_submit_packed -> b200scan_collect8, three
            slots in flight as in the CLI; every step packs, uploads, scores and downloads its hit list; host clock, max over ranks.
--config c3 (configs[2]): the drop-in CLI end to end, strong scaling: 3.1 Gbp in 24 manifest groups (GC 0.36 .. 0.48, N runs),
  `blamm-b200 dict / hist / scan -rc -pt 1e-4 -g N`, FASTA in -> occurrences.txt out.  e2e = wall clock of the scan process.
--config c4 (configs[3]): 10,000 PWMs of length 6-30 (20,000 columns) x 1 Gbp, `-at 12`, blocks of 100 Mbp dealt to the ranks.
--config c5 (configs[4]): `blamm-b200 hist -e` over the whole input on all GPUs, then `scan -pt` from those histograms.
--impl reference times the reference's own CPU implementation (oracle/_ref/blamm, the unmodified sources compiled by
  oracle/build_ref.sh; the C oracle port only if that binary is absent) with all host threads on a bounded sample of the same
  workload; that arm imports nothing of this repository's native code.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "window x motif x strand scores per second"
UNIT = "scores/s"
N_MOTIFS, MOTIF_SEED, SEQ_SEED = 900, 2024, 4242
SPECIES = "syn"
HUMAN_LIKE = (0.295, 0.205, 0.205, 0.295)          # background-biased variant (A/T 0.295, C/G 0.205)
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "blamm")


def workload_c2(n_gpus: int, mbp: float, bias: bool, softmask: float) -> dict:
    """`config` of both arms for c2 (identical text in the GPU arm and in the reference arm)."""
    return {"workload": "configs[1]: JASPAR-like %d PWMs x2 strands (1800 columns) x %d x %.0f Mbp synthetic %s ACGT "
                        "(one stream, one %.0f Mbp chunk per GPU), -rc -pt 1e-4 (real JASPAR CORE unavailable offline)%s"
                        % (N_MOTIFS, n_gpus, mbp, "background-biased (A/T 0.295, C/G 0.205)" if bias else "uniform", mbp,
                           (", %.0f %% soft-masked" % (100 * softmask)) if softmask > 0 else ""),
            "l2": "GPU arm: L2 flushed between timed steps (256 MiB memset outside the event pairs); every step reads a 100 Mbp chunk"}


def measured_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_filter_traffic.json")))
    if not files:
        return None
    try:
        return json.load(open(files[-1]))["traffic_bytes_per_launch"]
    except Exception:
        return None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1590.0), d.get("bf16_tflops_sustained", 1400.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons of one GPU, sampled DURING the timed region: an NVML thread
    (10 ms period, device addressed by PCI bus id so CUDA_VISIBLE_DEVICES cannot confuse it); if pynvml is unusable,
    the `nvidia-smi --query-gpu ... -lms` loop of the profiling recipe."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, pci_bus_id: str = None):
        self.idx, self.bus, self.f, self.p = gpu_index, pci_bus_id, None, None
        self.rows, self.thread, self.stop_flag, self.nv, self.h = [], None, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByPciBusId(pci_bus_id.encode()) if pci_bus_id else pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            pynvml.nvmlDeviceGetClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                try:
                    reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM), nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0, reasons))
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def start(self):
        if self.nv:
            import threading
            self.stop_flag = threading.Event()
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()
            return
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                       "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.nv:
            nv = self.nv
            self.stop_flag.set(); self.thread.join()
            if not self.rows:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
            busy = [r[0] for r in self.rows if r[1] > 200.0] or [r[0] for r in self.rows]
            bits = 0
            for r in self.rows:
                bits |= r[2]
            names = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                     ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            return {"sm_mhz": float(statistics.median(busy)), "sm_max_mhz": float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)),
                    "reasons": [n for n, b in names if bits & b], "samples": len(self.rows), "power_w_max": max(r[1] for r in self.rows),
                    "source": "nvml, 10 ms period"}
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in open(self.f.name).read().splitlines() if r.count(",") >= 8]
        os.unlink(self.f.name)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = [float(r[1]) for r in rows]
        busy = [s for s, r in zip(sm, rows) if float(r[3]) > 200.0] or sm
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[5 + k].strip() == "Active" for r in rows)]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": float(rows[0][2]), "reasons": reasons, "samples": len(rows),
                "power_w_max": max(float(r[3]) for r in rows), "source": "nvidia-smi -lms 100"}


def pci_bus_id(device_index: int) -> str:
    import torch
    p = torch.cuda.get_device_properties(device_index)
    try:
        return "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
    except Exception:
        return None


def bind_to_gpu_numa_node(bus: str) -> str:
    """Pin this rank (and so its pinned host buffers, first touch) to the CPUs of the GPU's NUMA node: with one process per
    GPU the hit downloads of all ranks otherwise meet on one socket's memory controllers."""
    try:
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus[-12:].lower()).read())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception as e:
        return "numa: unbound (%s)" % type(e).__name__
    return "numa: unbound"




def tensor_peaks(device: int):
    """(int8 Top/s, f16 TFLOP/s, source): the dense tcgen05.mma rate of the pipe the filter runs on, measured live on this box by
    blamm_b200/lib/i8_peak (tools/micro/i8_peak.cu: M128 N256 MMAs out of shared memory on every SM); else the figure committed
    under profiles/; else twice the bf16 number of MEASURED_PEAKS.json, labelled nominal."""
    import glob
    exe = os.path.join(ROOT, "blamm_b200", "lib", "i8_peak")
    try:
        out = subprocess.run([exe, "40000", str(device)], capture_output=True, text=True, timeout=120, check=True).stdout
        d = json.loads(out.strip().splitlines()[-1])
        return d["i8_tops"], d["f16_tflops"], "measured live by tools/micro/i8_peak.cu (dense tcgen05.mma kind::i8 loop, all SMs, SM clock %d MHz under that load)" % d["sm_mhz_i8"]
    except Exception:
        pass
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_i8_peak.json")), reverse=True):
        try:
            d = json.load(open(f))
            return d["i8_tops"], d["f16_tflops"], "measured earlier on this pool (%s)" % os.path.relpath(f, ROOT)
        except Exception:
            pass
    burst, _, how = peaks()
    return 2.0 * burst, burst, "NOMINAL: 2 x the bf16 cuBLAS burst, " + how


def make_motifs(workdir: str, config: str) -> str:
    from blamm_b200 import synth
    mfile = os.path.join(workdir, "motifs.jaspar")
    if config == "c4":
        synth.make_jaspar_like(mfile, 10000, 77, uniform_len=(6, 30))
    else:
        synth.make_jaspar_like(mfile, N_MOTIFS, MOTIF_SEED)
    return mfile


def build_inputs(workdir: str, n_nt: int, k: int = 0):
    """c2: motif set, P, thresholds (`-rc -pt 1e-4` from theoretical histograms, host C++ model) and chunk k of the stream -- the
    inputs run_blocks() scans; tests/test_gpu_parity.py checks the full-size workload through this function."""
    from blamm_b200 import capi, synth
    ms = capi.MotifSet(make_motifs(workdir, "c2"), revcompl=True)
    seq = chunk_chars(k, n_nt, (0.25, 0.25, 0.25, 0.25))
    first = seq[: min(n_nt, 4_000_000)]
    bg = [int(round(c * (n_nt / len(first)))) for c in synth.counts_of(first)]
    ms.write_histograms(bg, SPECIES, workdir)
    P, col_len, _ = ms.generate_matrix(bg)
    thr = ms.thresholds("pt", 1e-4, SPECIES, workdir)
    return ms, P, col_len, thr, seq, bg


def chunk_chars(k: int, n_nt: int, probs, softmask: float = 0.0) -> np.ndarray:
    """Chunk k of the synthetic stream (every chunk has its own seed, so a rank generates only what it scans)."""
    from blamm_b200 import synth
    seq = synth.random_acgt(n_nt, SEQ_SEED + k, probs)
    if softmask > 0:
        rng = np.random.default_rng(7 + k)
        p = 0
        while p < n_nt:
            run = int(rng.integers(1, 3000))
            if rng.random() < softmask:
                seq[p:p + run] |= 0x20
            p += run
    return seq


def write_ref_fasta(path: str, seq: np.ndarray, records: int = 4) -> int:
    from blamm_b200 import synth
    q = len(seq) // records
    synth.write_fasta(path, [("chr%d" % (i + 1), seq[i * q:(i + 1) * q]) for i in range(records)])
    return q * records


def ref_run(args_list, cwd, env) -> float:
    t0 = time.perf_counter()
    subprocess.run([REF_BIN] + args_list, cwd=cwd, env=env, check=True, stdout=subprocess.DEVNULL)
    return time.perf_counter() - t0


def reference_passes(work: str, scan_args, n_scores: float, warmup: int, steps: int, budget_s: float):
    """Timed passes of the reference `scan` over the prepared inputs in `work`: at most one warm pass, then as many timed
    passes as fit the budget (>= 1, <= steps).  Returns (times, cores)."""
    cores = os.cpu_count() or 1
    env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
    cmd = ["scan"] + scan_args + ["-t", str(cores), "motifs.jaspar", "sequences.mf"]
    est = None
    if warmup > 0:
        est = ref_run(cmd, work, env)
    times = [ref_run(cmd, work, env)]
    est = est or times[0]
    while len(times) < steps and (len(times) + 2) * est < budget_s:
        times.append(ref_run(cmd, work, env))
    return times, cores


def run_reference(args, rank: int, world: int) -> None:
    """The reference arm: the unmodified reference binary, `-t <all cores>`, OPENBLAS_NUM_THREADS=1, on a slice of the GPU arm's
    own input that keeps every core busy (>= 100 Mbp = 400 of the reference's 250,000-character blocks for c2/c3), plus one
    `-at 1000` pass (no occurrences: compute only).  Inputs come from blamm_b200.synth (numpy) and the
    reference's own `dict` / `hist`; nothing native of this repository is loaded."""
    if rank != 0:
        return
    from blamm_b200 import synth
    work = tempfile.mkdtemp(prefix="bench_ref_")
    try:
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
        make_motifs(work, args.config)
        probs = HUMAN_LIKE if args.bias else (0.25, 0.25, 0.25, 0.25)
        if args.config == "c4":
            n_nt = int(args.ref_mbp * 1e6) if args.ref_mbp else 4_000_000
            seq = synth.random_acgt(n_nt, SEQ_SEED, probs)
            scan_args, n_cols = ["-rc", "-at", "12"], 20000
            config = workload_c4(args.gpus, args.mbp if args.mbp else 1000.0)
        elif args.config in ("c3", "c5"):
            n_nt = int(args.ref_mbp * 1e6) if args.ref_mbp else 100_000_000
            gc = 0.36
            seq = synth.random_acgt(n_nt, 500, (0.5 - gc / 2, gc / 2, gc / 2, 0.5 - gc / 2))
            scan_args, n_cols = ["-rc", "-pt", "0.0001"], 2 * N_MOTIFS
            config = workload_c3(args.gpus, args.gbp) if args.config == "c3" else workload_c5(args.gpus, args.gbp)
        else:
            n_nt = int(args.ref_mbp * 1e6) if args.ref_mbp else int(min(args.mbp if args.mbp else 100.0, 100.0) * 1e6)
            seq = chunk_chars(0, n_nt, probs, args.softmask)          # the GPU arm's chunk 0
            scan_args, n_cols = ["-rc", "-pt", "0.0001"], 2 * N_MOTIFS
            config = workload_c2(args.gpus, args.mbp if args.mbp else 100.0, args.bias, args.softmask)
        usable = os.path.exists(REF_BIN)
        if usable:
            n_nt = write_ref_fasta(os.path.join(work, "sample.fa"), seq)
            open(os.path.join(work, "sequences.mf"), "w").write("%s\tsample.fa\n" % SPECIES)
            ref_run(["dict", "sequences.mf"], work, env)
            extra = {}
            if args.config == "c5":
                t_hist = ref_run(["hist", "-e", "-t", str(os.cpu_count() or 1), "motifs.jaspar", "sequences.mf"], work, env)
                extra["hist_e_s"] = t_hist
                extra["hist_e_note"] = "reference `hist -e` (default -l 10,000,000 characters of the group) on all cores"
            elif "-pt" in scan_args:
                ref_run(["hist", "motifs.jaspar", "sequences.mf"], work, env)
            times, cores = reference_passes(work, scan_args, n_nt * n_cols, min(args.warmup, 1), max(args.steps, 1), args.ref_budget)
            t_compute = ref_run(["scan", "-rc", "-at", "1000", "-t", str(cores), "motifs.jaspar", "sequences.mf"], work, env)
            extra["compute_only"] = {"value": n_nt * n_cols / t_compute, "unit": UNIT, "how": "`scan -rc -at 1000` (no occurrence passes the threshold), same slice, %.1f s" % t_compute}
            value = n_nt * n_cols * len(times) / sum(times)
            sample = ("%.1f Mbp slice (%d blocks of the reference's 250,000 characters) x %d columns per step, `blamm scan %s -t %d`, OPENBLAS_NUM_THREADS=1, "
                      "wall clock of the whole process; %d timed pass(es) after %d warm pass; scales linearly in sequence length"
                      % (n_nt / 1e6, n_nt // 250000, n_cols, " ".join(scan_args), cores, len(times), min(args.warmup, 1)))
            kind = "reference"
        else:                                                     # the checker's C port, single thread (bench.py may execute oracle/ here)
            from blamm_b200 import capi
            from oracle import oracle as O
            ms = capi.MotifSet(os.path.join(work, "motifs.jaspar"), revcompl=True)
            P, col_len, _ = ms.generate_matrix(synth.counts_of(seq[:1_000_000]))
            thr = np.full(len(col_len), 12.0, dtype=np.float32)
            n_nt, n_cols, cores, extra = 200_000, len(col_len), 1, {}
            times = []
            for _ in range(max(1, min(args.steps, 3))):
                t0 = time.perf_counter()
                O.scan_stream(bytes(seq[:n_nt]), np.zeros(1, np.uint64), P, col_len, thr)
                times.append(time.perf_counter() - t0)
            value = n_nt * n_cols * len(times) / sum(times)
            sample, kind = "%.2f Mbp x %d columns per step, single-thread C oracle port (oracle/_ref/blamm absent)" % (n_nt / 1e6, n_cols), "port"
        line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(times),
                "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
                "scaling": "strong" if args.config in ("c3", "c4", "c5") else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": dict({"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}, **extra),
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(work, ignore_errors=True)


def cpu_baseline(work: str, seq: np.ndarray, n_cols: int, sample_nt: int, scan_args, device=None) -> dict:
    """Reference binary on a bounded sample of the same workload (rank 0, N = 1 only; ~20 s of wall clock on 16 cores)."""
    cores = os.cpu_count() or 1
    sub = os.path.join(work, "cpu")
    os.makedirs(sub, exist_ok=True)
    try:
        if not os.path.exists(REF_BIN):
            raise RuntimeError("oracle/_ref/blamm absent")
        shutil.copy(os.path.join(work, "motifs.jaspar"), sub)
        n = write_ref_fasta(os.path.join(sub, "sample.fa"), seq[:sample_nt])
        open(os.path.join(sub, "sequences.mf"), "w").write("%s\tsample.fa\n" % SPECIES)
        env = dict(os.environ, OPENBLAS_NUM_THREADS="1")
        ref_run(["dict", "sequences.mf"], sub, env)
        if "-pt" in scan_args:
            ref_run(["hist", "motifs.jaspar", "sequences.mf"], sub, env)
        dt = ref_run(["scan"] + scan_args + ["-t", str(cores), "motifs.jaspar", "sequences.mf"], sub, env)
        out = {"value": n * n_cols / dt, "unit": UNIT, "cores": cores, "kind": "reference",
               "sample": "first %.1f Mbp of the workload (%d reference blocks) x %d columns, `blamm scan %s -t %d` (OpenBLAS threads=1), %.1f s wall" % (
                   n / 1e6, n // 250000, n_cols, " ".join(scan_args), cores, dt)}
        # the reference's OWN GPU path on this B200 (`scan -c`: cuBLAS sgemm per offset + filterScore, pwmscan.cpp:297-437, kernel.cu),
        # compiled for sm_100 by oracle/build_ref.sh: the existing GPU implementation next to the B200-native one
        ref_cuda = os.path.join(ROOT, "oracle", "_ref", "blamm_cuda")
        if os.path.exists(ref_cuda) and device is not None:
            try:
                env1 = dict(env, CUDA_VISIBLE_DEVICES=str(device))
                t0 = time.perf_counter()
                subprocess.run([ref_cuda, "scan", "-c"] + scan_args + ["-t", str(cores), "-o", "occ_cuda.txt", "motifs.jaspar", "sequences.mf"],
                               cwd=sub, env=env1, check=True, stdout=subprocess.DEVNULL, timeout=600)
                dtc = time.perf_counter() - t0
                n_cpu = sum(1 for _ in open(os.path.join(sub, "occurrences.txt"), "rb"))
                n_gpu = sum(1 for _ in open(os.path.join(sub, "occ_cuda.txt"), "rb"))
                out["gpu_reference"] = {"value": n * n_cols / dtc, "unit": UNIT, "how": "reference `blamm scan -c` (cuBLAS sgemm per offset + filterScore, unmodified sources "
                                        "compiled for sm_100) on 1 B200, same %.1f Mbp sample, %.1f s wall; occurrences: %d (its CPU path: %d)" % (n / 1e6, dtc, n_gpu, n_cpu)}
            except Exception as e:
                out["gpu_reference"] = {"value": None, "how": "reference `scan -c` failed: %s" % e}
        return out
    except Exception as e:                                        # reference binary unusable: time the C oracle port instead
        from blamm_b200 import capi, synth
        from oracle import oracle as O
        ms = capi.MotifSet(os.path.join(work, "motifs.jaspar"), revcompl=True)
        P, col_len, _ = ms.generate_matrix(synth.counts_of(seq[:1_000_000]))
        thr = np.full(len(col_len), 12.0, dtype=np.float32)
        n = 200_000
        t0 = time.perf_counter()
        O.scan_stream(bytes(seq[:n]), np.zeros(1, np.uint64), P, col_len, thr)
        dt = time.perf_counter() - t0
        return {"value": n * n_cols / dt, "unit": UNIT, "cores": 1, "kind": "port",
                "sample": "first %.2f Mbp x %d columns, single-thread C oracle (%s)" % (n / 1e6, n_cols, e)}


class HostPacker:
    """The CLI's hand-over in bench form: characters in pinned host memory -> 2-bit codes (+ zero mask) in pinned host memory by
    blamm_pack_ascii (libblammhost.so, SSE2) on `threads` host threads, 1 MiB of characters per task (ctypes releases the GIL)."""
    TASK = 1 << 20

    def __init__(self, threads: int):
        from concurrent.futures import ThreadPoolExecutor
        from blamm_b200 import capi
        self.L = capi.host_lib()
        self.threads = max(1, threads)
        self.pool = ThreadPoolExecutor(self.threads)
        self.driver = ThreadPoolExecutor(1)

    def _range(self, src, lo, hi, codes, zmask, fold):
        return self.L.blamm_pack_ascii(src + lo, hi - lo, fold, codes + (lo // 16) * 4, zmask + (lo // 32) * 4)

    def pack(self, src: int, n: int, codes: int, zmask: int, fold: int = 0) -> bool:
        t0 = time.perf_counter()
        futs = [self.pool.submit(self._range, src, lo, min(n, lo + self.TASK), codes, zmask, fold) for lo in range(0, n, self.TASK)]
        rcs = [f.result() for f in futs]
        if any(r < 0 for r in rcs):
            raise RuntimeError("blamm_pack_ascii failed")
        self.last_ms = 1e3 * (time.perf_counter() - t0)
        return any(r > 0 for r in rcs)

    def pack_async(self, *a):
        return self.driver.submit(self.pack, *a)

    def close(self):
        self.pool.shutdown(); self.driver.shutdown()


def workload_c4(n_gpus: int, mbp: float) -> dict:
    return {"workload": "configs[3]: 10,000 synthetic PWMs of length 6-30 x2 strands (20,000 columns) x %.0f Mbp synthetic uniform ACGT, -rc -at 12, "
                        "16 blocks dealt to %d GPU(s)" % (mbp, n_gpus),
            "l2": "GPU arm: every timed launch reads a different block's codes (16 MB at 1 Gbp) and a 5 MB weight image; the hit lists of a block (0.2-0.3 GB) exceed L2"}


def workload_c3(n_gpus: int, gbp: float) -> dict:
    return {"workload": "configs[2]: JASPAR-like %d PWMs x2 strands x %.2f Gbp synthetic genome in 24 manifest groups (GC 0.36 .. 0.48, N runs), "
                        "`dict`, `hist`, `scan -rc -pt 1e-4 -g %d`, FASTA in -> occurrences.txt out" % (N_MOTIFS, gbp, n_gpus),
            "l2": "GPU arm: every launch scores a different 32 Mi-character chunk (inputs + hit lists far larger than L2)"}


def workload_c5(n_gpus: int, gbp: float) -> dict:
    return {"workload": "configs[4]: JASPAR-like %d PWMs, `hist -e` (empirical score histograms over every whole group) then `scan -rc -pt 1e-4` with the cut-offs "
                        "read from them; %.2f Gbp in 24 manifest groups, -g %d" % (N_MOTIFS, gbp, n_gpus),
            "l2": "GPU arm: every launch scores a different chunk (inputs + hit lists far larger than L2)"}


def run_blocks(args, rank: int, local_rank: int, world: int) -> None:
    """c2 and c4: the hot path through the C ABI, one process per GPU."""
    import ctypes
    import torch
    import torch.distributed as dist
    from blamm_b200 import capi, shard, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the scan path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    bus = pci_bus_id(local_rank)
    numa = bind_to_gpu_numa_node(bus) if bus else "numa: unknown bus id"
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    c4 = args.config == "c4"
    # c4: the stream is always cut into 16 blocks (62.5 Mbp at the nominal 1 Gbp), so that 1, 2, 4 and 8 ranks get equal shares
    block_nt = int((args.mbp or 1000.0) * 1e6 / 16) if c4 else int((args.mbp or 100.0) * 1e6)
    n_blocks_total = 16 if c4 else world
    probs = HUMAN_LIKE if args.bias else (0.25, 0.25, 0.25, 0.25)
    work = tempfile.mkdtemp(prefix="bench_b200_")
    packer = None
    try:
        # ---- motif set, thresholds (host C++ model through its C ABI) ----
        mfile = make_motifs(work, args.config)
        ms = capi.MotifSet(mfile, revcompl=True)
        halo = ms.max_len - 1
        stream_len = block_nt * n_blocks_total
        plan = shard.plan_shards(stream_len, world, halo, block_nt)
        mine = [s for s in plan if s.rank == rank]
        first = chunk_chars(0, min(block_nt, 4_000_000), probs, args.softmask)
        bg = [int(round(c * (stream_len / len(first)))) for c in synth.counts_of(first)]     # i.i.d. synthetic sequence: counts scale linearly
        P, col_len, is_rc = ms.generate_matrix(bg)
        if c4:
            thr = ms.thresholds("at", 12.0)
            scan_args = ["-rc", "-at", "12"]
        else:
            ms.write_histograms(bg, SPECIES, work)
            thr = ms.thresholds("pt", 1e-4, SPECIES, work)
            scan_args = ["-rc", "-pt", "0.0001"]
        n_cols, sum_len = len(col_len), int(col_len.sum())

        # ---- the rank's chunks as characters in pinned host memory (chunk k = its own seeded block + the head of block k+1) ----
        L = capi.scan_lib()
        blocks = []
        seq0 = None
        for s in mine:
            chars = chunk_chars(s.index, s.n_payload, probs, args.softmask)
            if s.index == 0:
                seq0 = chars
            ptr = L.b200scan_host_alloc(s.n_total + 64)
            if not ptr:
                raise SystemExit("pinned allocation failed")
            ctypes.memmove(ptr, chars.ctypes.data, s.n_payload)
            if s.n_total > s.n_payload:
                nxt = chunk_chars(s.index + 1, 1 << 16, probs, 0.0)[: s.n_total - s.n_payload]
                ctypes.memmove(ptr + s.n_payload, nxt.ctypes.data, len(nxt))
            blocks.append((s, ptr))
            if s.index != 0:
                del chars
        max_total = max(s.n_total for s in mine)
        cw, zw = (max_total + 15) // 16, (max_total + 31) // 32
        n_code_bufs = 4
        code_bufs = [(L.b200scan_host_alloc(cw * 4 + 64), L.b200scan_host_alloc(zw * 4 + 64)) for _ in range(n_code_bufs)]
        cores = len(os.sched_getaffinity(0))
        packer = HostPacker(args.pack_threads or max(1, min(16, cores // world)))
        rate = 2.2e-4 if not c4 else 2.5e-5
        sc = capi.Scanner(local_rank, max_block_nt=max_total + 64, max_hits=max(1 << 20, int(rate * max_total * n_cols)))
        sc.set_engine({"auto": capi.ENGINE_AUTO, "tensor": capi.ENGINE_TENSOR, "gather": capi.ENGINE_GATHER}[args.engine])
        if args.acc:
            sc.set_tensor_accumulator(args.acc)
        sc.set_hit_format(args.hits)
        sc.set_motifs(P, col_len, thr)
        scores_per_step = sum(s.n_payload for s in mine) * n_cols
        total_scores_per_step = stream_len * n_cols
        ordered = args.hits == capi.HITS_8

        def submit(slot, k, buf, ascii_mode=None):
            s, ptr = blocks[k]
            if args.ascii if ascii_mode is None else ascii_mode:
                sc.submit_ascii(slot, ptr, n_total=s.n_total, n_payload=s.n_payload)
            else:
                sc.submit_packed_ptr(slot, buf[0], buf[1] if has_zero[k] else None, s.n_total, s.n_payload)

        def collect(slot):
            if ordered:
                h, b, t = sc.collect8(slot, copy=False)
                return len(h), len(h) * 8 + len(b) * 4 + 32, t, (h, b)
            h, t = sc.collect(slot, copy=False)
            return len(h), len(h) * args.hits + 32, t, (h, None)

        # ---- warm-up (also tells which chunks carry a zero mask) ----
        has_zero = [False] * len(blocks)
        n_hits = [0] * len(blocks)
        for w in range(args.warmup):
            for k, (s, ptr) in enumerate(blocks):
                if not args.ascii:
                    has_zero[k] = packer.pack(ptr, s.n_total, code_bufs[0][0], code_bufs[0][1])
                submit(w % capi.NUM_SLOTS, k, code_bufs[0])             # every slot allocates its buffers at its first use: outside the timed region
                n_hits[k], _, t_last, _ = collect(w % capi.NUM_SLOTS)

        sampler = ClockSampler(local_rank, bus)
        # ---- timed: device-resident steps (CUDA events on the scan stream), L2 flushed between launches ----
        barrier()
        sampler.start()
        tot_ms = k_ms = 0.0
        for step in range(args.steps):
            for k in range(len(blocks)):
                if len(blocks) > 1:                                   # make block k resident (untimed upload); c2 keeps its one block
                    if not args.ascii:
                        packer.pack(blocks[k][1], blocks[k][0].n_total, code_bufs[0][0], code_bufs[0][1])
                    submit(0, k, code_bufs[0])
                    collect(0)
                sc.flush_l2()
                a, b, nh = sc.rerun_resident(0, 1)
                tot_ms += a; k_ms += b
                assert nh == n_hits[k] or os.environ.get("B200_BENCH_DIAG"), "resident re-run changed the hit count"     # DIAG: knock-out builds (tools/variants.sh)
        barrier()
        # ---- timed: end to end from characters in pinned host memory, the way the CLI drives the library: pack on the host
        #      threads, three slots in flight (chunk j+1 uploads and chunk j scores while the hits of chunk j-1 come down) ----
        nS = capi.NUM_SLOTS

        def e2e_pass(n_steps, ascii_mode):
            """n_steps passes over the rank's chunks, pipelined; returns (seconds, h2d bytes, d2h bytes, host pack ms, hits, last timing)."""
            order = [k for _ in range(n_steps) for k in range(len(blocks))]
            t_start = time.perf_counter()
            d2h_b = h2d_b = 0
            p_ms = 0.0
            hits_seen = 0
            t_last = None
            fut = None if ascii_mode else packer.pack_async(blocks[order[0]][1], blocks[order[0]][0].n_total, code_bufs[0][0], code_bufs[0][1])
            for j, k in enumerate(order):
                buf = code_bufs[j % n_code_bufs]
                if fut is not None:
                    fut.result(); p_ms += packer.last_ms
                if not ascii_mode and j + 1 < len(order):            # buffer (j+1) % 4 was last read by chunk j-3, collected in iteration j-1
                    k2 = order[j + 1]; b2 = code_bufs[(j + 1) % n_code_bufs]
                    fut = packer.pack_async(blocks[k2][1], blocks[k2][0].n_total, b2[0], b2[1])
                else:
                    fut = None
                submit(j % nS, k, buf, ascii_mode)
                s = blocks[k][0]
                h2d_b += s.n_total if ascii_mode else ((s.n_total + 15) // 16) * 4 + (((s.n_total + 31) // 32) * 4 if has_zero[k] else 0)
                if j >= nS - 1:
                    nh, nbytes, t_last, _ = collect((j - (nS - 1)) % nS)
                    d2h_b += nbytes; hits_seen += nh
            for j in range(max(0, len(order) - (nS - 1)), len(order)):
                nh, nbytes, t_last, _ = collect(j % nS)
                d2h_b += nbytes; hits_seen += nh
            torch.cuda.synchronize()
            return time.perf_counter() - t_start, h2d_b, d2h_b, p_ms, hits_seen, t_last

        # Hand-over: the ABI takes a block as characters (device packer: 1 B per character over PCIe, no host work) or as host-packed
        # 2-bit codes (0.25 B per character, but the host cores must keep up: N ranks share them).  Unless one is forced, both are
        # tried for a few untimed steps and the faster one (max over ranks) is used -- the choice a caller of the ABI would make.
        calib = None
        ascii_mode = args.ascii
        if not args.ascii and not args.packed:
            cal = []
            for mode in (False, True):
                e2e_pass(1, mode)                                     # first use of this hand-over's buffers
                barrier()
                cal.append(e2e_pass(max(4 // len(blocks), 1), mode)[0])
                barrier()
            ct = torch.tensor(cal, dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(ct, op=dist.ReduceOp.MAX)
            cal = ct.tolist()
            ascii_mode = cal[1] < 0.95 * cal[0]
            calib = {"packed_s": cal[0], "characters_s": cal[1], "steps": max(4 // len(blocks), 1)}
        barrier()
        e2e_s, h2d, d2h, pack_ms, checksum, t_e2e = e2e_pass(args.steps, ascii_mode)
        barrier()
        clocks = sampler.stop()
        assert checksum == sum(n_hits) * args.steps or os.environ.get("B200_BENCH_DIAG"), "e2e passes changed the hit count"

        stats = torch.tensor([tot_ms, k_ms, e2e_s * 1e3], dtype=torch.float64, device="cuda")
        counts = torch.tensor([sum(n_hits), sum(s.n_payload for s in mine)], dtype=torch.int64, device="cuda")
        all_counts = [counts]
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.MAX)
            all_counts = [torch.zeros_like(counts) for _ in range(world)]
            dist.all_gather(all_counts, counts)              # the stream-order merge needs only the per-rank hit counts
        tot_ms, k_ms, e2e_ms = stats.tolist()
        hits_total = int(sum(int(c[0]) for c in all_counts))
        launches_per_pass = t_e2e["kernel_launches"]
        engine_used = "tensor+rescore" if t_e2e["engine_used"] == capi.ENGINE_TENSOR else "gather"
        kinds = {8: "int8 operands, s32 accumulators (tcgen05.mma.kind::i8)", 16: "fp16 operands, fp16 accumulators (kind::f16)",
                 32: "fp16 operands, fp32 accumulators (kind::f16)", 0: "mixed per column tile (int8 / fp16)"}
        operands = kinds.get(sc.tensor_info()["accumulator_bits"], "?") if engine_used != "gather" else "fp32 gather-add"
        int8_pipe = engine_used != "gather" and sc.tensor_info()["accumulator_bits"] in (8, 0)
        tw = sc.tensor_work()

        if rank == 0:
            i8_peak, f16_peak, peak_how = tensor_peaks(local_rank)
            burst, sustained, how = peaks()
            launches = args.steps * len(blocks)
            per_launch_nt = sum(s.n_payload for s in mine) / len(mine)
            flops_per_launch = tw["algorithmic_ops_per_window"] * per_launch_nt          # 8 x sum L per window, un-padded
            mma_per_launch = tw["mma_ops_per_window"] * per_launch_nt
            kernel_ms = k_ms / launches
            achieved = flops_per_launch / (kernel_ms * 1e-3) / 1e12
            peak = i8_peak if int8_pipe else f16_peak
            config = workload_c4(world, args.mbp or 1000.0) if c4 else workload_c2(world, args.mbp or 100.0, args.bias, args.softmask)
            line = {
                "metric": METRIC, "value": total_scores_per_step * args.steps / (tot_ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": tot_ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if c4 else "weak", "vs_baseline": None, "dtype": "s8 filter / f32 scores" if int8_pipe else "f32", "data": "synthetic",
                "config": config,
                "path": {"engine": engine_used, "operands": operands, "columns": n_cols, "sum_len": sum_len,
                         "parallelism": "one stream of %d chunk(s) dealt to %d rank(s) (chunk k -> rank k mod N, halo %d), no data-path collective; "
                                        "merge = chunk-ordered concatenation of per-chunk (position, column)-ordered hit lists, ranks exchange hit counts only" % (len(plan), world, halo),
                         "host_binding": numa, "hits_per_step": hits_total, "hit_record_bytes": args.hits,
                         "candidates_last_block": int(t_e2e["n_candidates"]), "hand_over": "characters in pinned host memory (b200scan_submit_ascii: 1 B per character up, packed on the device)" if ascii_mode else
                         "host 2-bit packer on %d threads (blamm_pack_ascii) + b200scan_submit_packed" % packer.threads,
                         "hand_over_calibration": calib,
                         "slots_in_flight": nS},
                "gpu_launches": int(launches * (2 * launches_per_pass - (1 if ascii_mode else 0))),     # (the character hand-over adds the pack kernel to the e2e passes)
                "clocks": clocks,
                "e2e": {"value": total_scores_per_step * args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(h2d // args.steps),
                        "d2h_bytes_per_step": int(d2h // args.steps), "ms_per_step": e2e_ms / args.steps,
                        "host_pack_ms_per_step": pack_ms / args.steps,
                        "stages_ms_last_block": {k: t_e2e[k] for k in ("h2d_ms", "pack_ms", "score_ms", "rescore_ms", "order_ms", "d2h_ms")}},
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "pipe": {"achieved": mma_per_launch / (kernel_ms * 1e-3) / 1e12, "frac": mma_per_launch / (kernel_ms * 1e-3) / 1e12 / peak,
                                      "what": "tensor-core operations as issued: column tiles padded to N and to whole K steps (b200scan_tensor_work)"},
                             "traffic": measured_traffic() if (engine_used != "gather" and not c4 and abs((args.mbp or 100.0) - 100.0) < 1e-9) else None,
                             "peak_source": ("%s of the pipe this kernel issues to; " % ("INT8 (kind::i8)" if int8_pipe else "FP16 (kind::f16)")) + peak_how,
                             "peak_bf16_cublas": {"burst": burst, "sustained": sustained, "source": how},
                             "kernel": "filter_tc_kernel" if engine_used != "gather" else "gather_scan_kernel",
                             "kernel_ms": kernel_ms, "algorithmic_flops_per_launch": flops_per_launch, "mma_flops_per_launch": mma_per_launch,
                             "note": "achieved / frac = ALGORITHMIC operations (8 x sum L per window, integer multiply-adds counted like flops) per measured kernel time"},
            }
            if world == 1 and not args.no_cpu_baseline and seq0 is not None:
                line["cpu_baseline"] = cpu_baseline(work, seq0, n_cols, min(args.cpu_sample_nt if not c4 else 4_000_000, len(seq0)), scan_args, local_rank)
            print(json.dumps(line), flush=True)
        sc.close()
        for _, ptr in blocks:
            L.b200scan_host_free(ptr)
        for a, b in code_bufs:
            L.b200scan_host_free(a); L.b200scan_host_free(b)
    finally:
        if packer:
            packer.close()
        shutil.rmtree(work, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def run_cli(args, rank: int, local_rank: int, world: int) -> None:
    """c3 and c5: the drop-in command line end to end (strong scaling: one input, -g N).  Under torchrun the ranks only share
    the generation of the FASTA set (gloo; no GPU work in this process): rank 0 then runs `blamm-b200 ... -g N` and reports."""
    import torch.distributed as dist
    from blamm_b200 import synth
    cli = os.path.join(ROOT, "blamm_b200", "lib", "blamm-b200")
    if not os.path.exists(cli):
        raise SystemExit("bench.py: blamm_b200/lib/blamm-b200 is missing (run `make`)")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    n_groups = 24
    per_group = int(args.gbp * 1e9 / n_groups)
    box = [None]
    reuse = bool(args.reuse) and os.path.exists(os.path.join(args.reuse, "sequences.mf"))
    if rank == 0 and args.reuse:
        os.makedirs(args.reuse, exist_ok=True)
        box[0] = args.reuse
    elif rank == 0:
        base = args.workdir
        if not base:
            try:
                st = os.statvfs("/dev/shm")
                base = "/dev/shm" if st.f_bavail * st.f_frsize > 20 * args.gbp * 1e9 + (8 << 30) else tempfile.gettempdir()
            except Exception:
                base = tempfile.gettempdir()
        box[0] = tempfile.mkdtemp(prefix="bench_cli_", dir=base)
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    work = box[0]
    try:
        t_gen = time.perf_counter()
        for g in ([] if reuse else range(rank, n_groups, world)):
            gc = 0.36 + 0.12 * g / (n_groups - 1)
            seq = synth.random_acgt(per_group, 500 + g, (0.5 - gc / 2, gc / 2, gc / 2, 0.5 - gc / 2))
            rng = np.random.default_rng(900 + g)
            for _ in range(6):                                       # a few N runs (assembly gaps): they split records into fragments
                a = int(rng.integers(0, max(1, per_group - 60000)))
                seq[a:a + int(rng.integers(1, 50000))] = ord("N")
            q = per_group // 3
            synth.write_fasta(os.path.join(work, "g%02d.fa" % g), [("g%02d_chr%d" % (g, i + 1), seq[i * q:(i + 1) * q]) for i in range(3)])
            del seq
        if world > 1:
            dist.barrier()
        t_gen = time.perf_counter() - t_gen
        if rank == 0:
            cores = os.cpu_count() or 1
            with open(os.path.join(work, "sequences.mf"), "w") as mf:
                for g in range(n_groups):
                    mf.write("group%02d\tg%02d.fa\n" % (g, g))
            make_motifs(work, "c3")
            env = dict(os.environ)

            def run(cmd):
                t0 = time.perf_counter()
                r = subprocess.run([cli] + cmd, cwd=work, env=env, capture_output=True, text=True)
                if r.returncode != 0:
                    raise SystemExit("bench.py: `blamm-b200 %s` failed:\n%s\n%s" % (" ".join(cmd), r.stdout[-2000:], r.stderr[-2000:]))
                return time.perf_counter() - t0, r.stdout
            t_dict, _ = run(["dict", "sequences.mf"])
            if args.config == "c5":
                t_hist, _ = run(["hist", "-e", "-l", str(per_group + 1), "-t", str(cores), "-g", str(world), "motifs.jaspar", "sequences.mf"])
            else:
                t_hist, _ = run(["hist", "-t", str(cores), "motifs.jaspar", "sequences.mf"])
            warm, steps = min(args.warmup, 1), max(1, min(args.steps, 3))
            sampler = ClockSampler(0, None)
            walls, stats = [], None
            out = os.path.join(work, "occurrences.txt")
            for p in range(warm + steps):
                if p == warm:
                    sampler.start()
                dt, text = run(["scan", "-rc", "-pt", "0.0001", "-g", str(world), "-t", str(cores), "-o", "occurrences.txt", "--stats", "stats.json",
                                "motifs.jaspar", "sequences.mf"])
                out_bytes = os.path.getsize(out)
                os.remove(out)
                if p >= warm:
                    walls.append(dt)
                    stats = json.load(open(os.path.join(work, "stats.json")))
            clocks = sampler.stop()
            n_cols = stats["columns"]
            chars = sum(d["characters"] for d in stats["devices"])
            scores = chars * n_cols
            kern = [d["score_ms"] + d["rescore_ms"] + d["order_ms"] for d in stats["devices"]]
            wall = sum(walls) / len(walls)
            launches = sum(d["chunks"] for d in stats["devices"])
            line = {"metric": METRIC, "value": scores / (max(kern) * 1e-3), "unit": UNIT, "n_gpus": world, "steps": len(walls), "warmup": warm,
                    "ms_per_step": max(kern), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "s8 filter / f32 scores",
                    "data": "synthetic", "config": workload_c3(world, args.gbp) if args.config == "c3" else workload_c5(world, args.gbp),
                    "path": {"what": "value = scores / the busiest GPU's summed kernel time (filter + expand + rescore + ordering, CUDA events per chunk, "
                                     "`scan --stats`); e2e = scores / wall clock of the whole `blamm-b200 scan` process (FASTA in -> occurrences.txt out)",
                             "filtered_characters": chars, "columns": n_cols, "matches": stats["matches"], "occurrences_bytes": out_bytes,
                             "host_cores": cores, "workdir": os.path.dirname(work), "dict_s": t_dict, "hist_s": t_hist,
                             "hist": "`hist -e` over every whole group on %d GPU(s)" % world if args.config == "c5" else "theoretical spectra (host)",
                             "generate_fasta_s": t_gen, "per_gpu_kernel_ms": kern, "phases_s": stats["phases_s"], "devices": stats["devices"]},
                    "gpu_launches": int(launches * 6 * len(walls)),
                    "clocks": clocks,
                    "e2e": {"value": scores / wall, "unit": UNIT, "wall_s": wall, "h2d_bytes_per_step": int(chars * 0.25),
                            "d2h_bytes_per_step": int(stats["matches"] * 8 + chars / 64)},
                    "roofline": None}
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.barrier()
    finally:
        if rank == 0 and not args.keep and not args.reuse:
            shutil.rmtree(work, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4", "c5"], help="c2 = BASELINE.json configs[1] (default), c3 = configs[2] (CLI end to end), c4 = configs[3], c5 = configs[4]")
    ap.add_argument("--mbp", type=float, default=0.0, help="c2: Mbp per GPU (default 100); c4: Mbp in total (default 1000)")
    ap.add_argument("--gbp", type=float, default=3.1, help="c3 / c5: size of the synthetic genome in Gbp")
    ap.add_argument("--bias", action="store_true", help="background-biased sequence (A/T 0.295, C/G 0.205) instead of uniform")
    ap.add_argument("--engine", default="auto", choices=["auto", "tensor", "gather"])
    ap.add_argument("--acc", type=int, default=0, choices=[0, 8, 16, 32], help="tensor filter operand/accumulator kind: 0 = automatic (diagnostic)")
    ap.add_argument("--softmask", type=float, default=0.0, help="diagnostic: fraction of the sequence turned lower case (runs of 1..3000), scored with BLAS-path semantics")
    ap.add_argument("--hits", type=int, default=8, choices=[8, 12, 16], help="hit record format (b200scan_set_hit_format): 8 = ordered b200scan_hit8, what the CLI uses")
    ap.add_argument("--ascii", action="store_true", help="force the character hand-over (b200scan_submit_ascii, device packer)")
    ap.add_argument("--packed", action="store_true", help="force the host-packed hand-over (blamm_pack_ascii + b200scan_submit_packed); default: the faster of the two")
    ap.add_argument("--pack-threads", type=int, default=0, help="host packer threads per rank (default: cores / ranks, at most 16)")
    ap.add_argument("--cpu-sample-nt", type=int, default=32_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-mbp", type=float, default=0.0, help="reference arm: size of the slice per pass (default 100 Mbp; c4: 4 Mbp x 20,000 columns)")
    ap.add_argument("--ref-budget", type=float, default=200.0, help="reference arm: seconds available for the timed passes")
    ap.add_argument("--workdir", default="", help="c3 / c5: directory for the FASTA set and the outputs (default: /dev/shm if roomy, else the temp dir)")
    ap.add_argument("--keep", action="store_true", help="c3 / c5: keep the work directory")
    ap.add_argument("--reuse", default="", help="c3 / c5: directory whose FASTA set is generated once and kept for later runs (e.g. the same input at -g 8 and -g 1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if (args.impl == "b200" and args.config in ("c2", "c4")) else args.warmup     # >= 3: one per slot

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.config in ("c2", "c4"):
        run_blocks(args, rank, local_rank, world)
    else:
        run_cli(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
