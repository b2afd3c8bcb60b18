#!/usr/bin/env python
"""bench.py — motif insertion-site search throughput on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: every unit of this
rank's library shard searched on the query (count -> scan -> emit kernels, queued without a
host round trip and replayed as a CUDA graph) and, when N > 1, the all-gather-v of the site
lists (padded single collective + device cut). One JSON line is printed by rank 0.

Workloads (BASELINE.json configs):
  c4        2000-nt query x 50 000 multi-strand JSON motifs per GPU (half 2-4 strands of 5-8 nt,
            half 3 strands of 3-4 nt), banded canonical pair mask thinned by Bernoulli(0.3)
            [default: BASELINE config "2000-nt x 50k motifs, search stress"; the sharded sweep
            1/2/4/8 GPUs runs it per rank = weak scaling]
  c4a / c4b the two halves on their own (search-bound / enumeration-bound)
  c4b_all   c4b with the all-allowed mask: every placement survives, output-bound
  c3        10 000 x 120-nt queries x 1 000-motif JSON library (batch mode)
  c2        230-nt query x 400 synthetic CaRNAval RINs at SURVEY 8d's shape (2-6 strands of 1-6 nt)
  c2_narrow the round-1 variant (2-4 strands of 2-6 nt), kept for comparison
  c1        73-nt example query x 1 000-motif JSON library
  c4_400k   c4 with a 400 000-module library (strong scaling with enough work per GPU)
  c4_desc   2000-nt query x 5 000 synthetic .desc modules (wildcards, widened strands, delay 5, gates)
Scaling (--gpus N > 1): "strong" (default) = ONE library, cut into N contiguous module ranges balanced by a
search-cost estimate (the reference's unit of parallelism is one job per module of one library,
cppsrc/MOIP.cpp:246-293); "weak" = one library of the named size PER GPU. The strong line carries the weak
number beside it, and the per-rank search / exchange / skew times.
value     = motif x position placements tested per second, inputs resident in HBM, device time
            (CUDA events on the launching stream), max over ranks.
e2e       = the same through the host-buffer C ABI call (bss_search*): upload of sequence + mask
            and download of the site records inside the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "motif_x_position_placements_tested_per_s"
UNIT = "placements/s"


# ----------------------------------------------------------------------------------------
# workloads
# ----------------------------------------------------------------------------------------
def make_workload(name, rank):
    """Returns dict(seqs, masks, modules | rin | desc, source, describe). `rank` seeds the library (weak scaling:
    rank r searches its own library of the named size; strong scaling and N = 1: the library of rank 0)."""
    from biorseo_b200 import synth
    w = {"name": name, "source": "jsonfolder"}
    if name in ("c4", "c4a", "c4b", "c4b_all", "c4_400k"):
        seq = synth.sequence(2000, 4)
        kind = {"c4": "mixed", "c4a": "c4a", "c4b": "c4b", "c4b_all": "c4b", "c4_400k": "mixed"}[name]
        size = 400000 if name == "c4_400k" else 50000
        w["seqs"] = [seq]
        w["masks"] = [synth.mask_all(2000) if name == "c4b_all" else synth.mask_canonical(seq, seed=204, keep=0.3, span=100)]
        w["modules"] = synth.json_modules(size, 104 + 1000 * rank, kind)
        w["describe"] = f"2000-nt x {size} JSON motifs ({kind}), mask=" + ("all-allowed" if name == "c4b_all" else "banded canonical, Bernoulli 0.3")
    elif name == "c3":
        n_seqs = 10000
        w["seqs"] = [synth.sequence(120, 3 + i) for i in range(n_seqs)]
        rng = np.random.default_rng(203)
        w["masks"] = [synth.mask_canonical(s, seed=int(rng.integers(1 << 30)), keep=0.3) for s in w["seqs"]]
        w["modules"] = synth.json_modules(1000, 101, "c1")
        w["describe"] = "10000 x 120-nt queries x 1000 JSON motifs (c1 library), banded canonical masks"
    elif name in ("c2", "c2_narrow"):
        seq = synth.sequence(230, 2)
        w["seqs"], w["masks"] = [seq], [synth.mask_canonical(seq, seed=202, keep=0.3)]
        w["source"] = "rinfolder"
        # (n, seed, max_components, min_len): SURVEY 8d's shape is 2-6 strands of 1-6 nt (scripts/Install_CaRNAval_RINs.py:54-101);
        # its one-nucleotide strands are where modules with 10^6-10^8 placements before the filters come from
        w["rin"] = (400, 102, 6, 1) if name == "c2" else (400, 102, 4, 2)
        w["describe"] = ("230-nt query x 400 synthetic CaRNAval RINs, 2-6 strands of 1-6 nt" if name == "c2" else
                         "230-nt query x 400 synthetic CaRNAval RINs, narrowed to 2-4 strands of 2-6 nt (round-1 shape)")
    elif name == "c4_desc":
        seq = synth.sequence(2000, 4)
        w["seqs"], w["masks"] = [seq], [synth.mask_canonical(seq, seed=204, keep=0.3, span=100)]
        w["source"] = "descfolder"
        w["desc"] = (5000, 11)
        w["describe"] = "2000-nt query x 5000 synthetic .desc modules (wildcards from gap-fill, widened short strands, delay 5, regex gates)"
    elif name == "c1":
        seq = "GGGGUAUCGCCAAGCGGUAAGGCACCGGAUUCUGAUUCCGGAGGUCGAGGUUCGAAUCCUCGUACCCCAGCCA"   # data/fasta/example.fa
        w["seqs"], w["masks"] = [seq], [synth.mask_canonical(seq, seed=201, keep=0.3)]
        w["modules"] = synth.json_modules(1000, 101, "c1")
        w["describe"] = "73-nt example.fa query x 1000 JSON motifs (c1 library)"
    else:
        raise SystemExit(f"unknown workload {name}")
    return w


def host_library(w, tmp):
    import biorseo_b200 as b
    from biorseo_b200 import synth
    if w["source"] == "rinfolder":
        n, seed, max_components, min_len = w["rin"]
        path = synth.write_rin_folder(os.path.join(tmp, "rin"), n, seed, invalid_fraction=0.0, max_components=max_components, min_len=min_len)
        return b.HostLibrary.load("rinfolder", path), path
    if w["source"] == "descfolder":
        n, seed = w["desc"]
        path = synth.write_desc_folder(os.path.join(tmp, "desc"), n, seed)
        return b.HostLibrary.load("descfolder", path), path
    return b.HostLibrary.from_json_modules(w["modules"]), None


def algorithmic_bytes(table, seq_lens, n_words):
    """B_in = per-unit table bytes + per-query (packed sequence + pair mask),
    B_out = 4 bytes per emitted word (u32 module + u32 start per component)."""
    n_units = int(table["n_units"])
    ucb = table["unit_comp_begin"].astype(np.int64)
    cpb = table["comp_pat_begin"].astype(np.int64)
    k = np.diff(cpb)[: int(ucb[n_units])]
    checks = int(table["unit_check_begin"][n_units])
    b_in = 4 * n_units + int((4 + (k + 3) // 4 + (k + 7) // 8).sum()) + 8 * checks
    b_in += int(sum((n + 3) // 4 + (n * n + 7) // 8 for n in seq_lens))
    return b_in, 4 * int(n_words)


def algorithmic_int_ops(table, seq_lens, n_sites):
    """SURVEY.md 8(d), integer regime: one funnel shift + one AND per pattern byte per 32-position word of
    every (query, unit, component) — the shift-AND lower bound of the match phase — plus 2 per placement
    (chain step + pair test); the emitted sites stand in for the placements (a lower bound of that term)."""
    n_units = int(table["n_units"])
    ucb = table["unit_comp_begin"].astype(np.int64)
    k = np.diff(table["comp_pat_begin"].astype(np.int64))[: int(ucb[n_units])]
    return int(2 * int(k.sum()) * sum((n + 31) // 32 for n in seq_lens) + 2 * int(n_sites))


# ----------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock + throttle reasons of this rank's GPU, sampled DURING the timed region. In-process NVML
    (nvidia_ml_py) when available: a query costs microseconds, where spawning nvidia-smi several times
    a second from every rank disturbs the very thing being timed; nvidia-smi is the fallback."""
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], threading.Event()
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag.is_set():
            try:
                if self.nvml is not None:
                    mhz = float(self.nvml.nvmlDeviceGetClockInfo(self.handle, self.nvml.NVML_CLOCK_SM))
                    try:
                        mask = int(self.nvml.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                    except Exception:
                        mask = int(self.nvml.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                    self.samples.append([mhz, self.max_mhz] + ["Active" if mask & bit else "Not Active" for _, bit in self.REASONS])
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.002 if self.nvml is not None else 0.2)

    def summary(self):
        self.stop_flag.set()
        self.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no clock samples (NVML and nvidia-smi unavailable)"]}
        sm = sorted(float(s[0]) for s in self.samples)
        names = [n for n, _ in self.REASONS]
        reasons = sorted({names[i] for s in self.samples for i in range(4) if len(s) >= 6 and str(s[2 + i]).lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons, "samples": len(sm),
                "source": "nvml" if self.nvml is not None else "nvidia-smi"}


# ----------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own Pool-threaded search (oracle/_ref), or
# the C++ port of it when the compiled reference is absent
# ----------------------------------------------------------------------------------------
def rin_placements_estimate(path, n):
    """Expected placements of a CaRNAval RIN file before the pair filters on a uniform n-nt query: the reference
    materialises every one of them (cppsrc/Motif.cpp:466-474), so modules with 10^6 and more cannot be run on the CPU."""
    with open(path) as f:
        lines = f.read().split("\n")[3:]
    est = 1.0
    for depth, line in enumerate(x for x in lines if x):
        k = len(line.split(";")[2])
        est *= max(1.0, n / 4.0 ** k) / (depth + 1)      # ordered placements of the strands
    return est


def write_sample(w, tmp, sample_modules):
    """The bounded sample of the workload's library the CPU arms run on (the reference's own on-disk formats).
    Returns (library path, what the sample is)."""
    from biorseo_b200 import synth
    if w["source"] == "rinfolder":
        n, seed, max_components, min_len = w["rin"]
        lib = synth.write_rin_folder(os.path.join(tmp, "rin"), n, seed, invalid_fraction=0.0, max_components=max_components, min_len=min_len)
        sub, kept, heavy = os.path.join(lib, "Subfiles"), 0, 0
        for i in range(n):     # module ids come from the file names: dropping files keeps the others' ids
            f = os.path.join(sub, f"{i}.txt")
            if kept >= sample_modules or rin_placements_estimate(f, len(w["seqs"][0])) > 2e5:
                heavy += kept < sample_modules
                os.remove(f)
            else:
                kept += 1
        return lib, f"{kept} of the {n} RIN modules (the {heavy} left out are predicted to have more than 2e5 placements before the filters, which the reference materialises)"
    if w["source"] == "descfolder":
        n, seed = w["desc"]
        m = min(n, sample_modules)
        return synth.write_desc_folder(os.path.join(tmp, "desc"), m, seed), f"first {m} of the {n} .desc modules"
    mods = w["modules"][:sample_modules]
    return synth.write_json(os.path.join(tmp, "lib.json"), mods), f"first {len(mods)} of the {len(w['modules'])} JSON modules"


def cpu_search(w, sample_modules, repeat, threads=None):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    kind = "reference" if util.have_ref() else "port"
    binary = util.REF_BIN if kind == "reference" else util.PORT_BIN
    if not os.access(binary, os.X_OK):
        return None
    cores = os.cpu_count() or 1
    threads = threads or max(1, cores - 1)          # the reference's own choice, cppsrc/MOIP.cpp:250
    seq, mask = w["seqs"][0], w["masks"][0]
    placements = None
    with tempfile.TemporaryDirectory() as tmp:
        lib, what = write_sample(w, tmp, sample_modules)
        seq_path, mask_path = util.write_query(tmp, seq, mask)
        t0 = time.perf_counter()
        if kind == "reference":
            out = subprocess.run([binary, "bench", w["source"], lib, seq_path, mask_path, str(threads), str(repeat)],
                                 capture_output=True, text=True, check=True).stdout
            res = json.loads(out.strip().splitlines()[-1])
            seconds, sites, n_modules = res["seconds_mean"], res["sites"], res["modules"]
            if w["source"] == "jsonfolder":   # the reference's real work term on the same sample (its own enumerator, untimed)
                placements = int(subprocess.run([binary, "placements", w["source"], lib, seq_path, mask_path], capture_output=True, text=True,
                                                check=True).stdout.strip().splitlines()[-1])
        else:   # single-threaded port, whole process timed
            threads = 1
            t1 = time.perf_counter()
            for _ in range(repeat):
                out = subprocess.run([binary, "jobs", w["source"], lib, seq_path, mask_path], capture_output=True, text=True, check=True).stdout
            seconds, sites = (time.perf_counter() - t1) / repeat, out.count("\n")
            n_modules = sample_modules
        wall = time.perf_counter() - t0
    tested = n_modules * len(seq) * 1   # one query of the workload
    return {"value": tested / seconds, "unit": UNIT, "cores": threads, "host_cores": cores, "kind": kind,
            "sample": f"{what} ({n_modules} valid) x query 0 ({len(seq)} nt), {repeat} pass(es), "
                      f"push->join timed ({seconds:.3f} s/pass, {wall:.1f} s total)",
            "sample_modules": n_modules, "sites_per_s": sites / seconds, "seconds_per_pass": seconds, "placements_enumerated": placements,
            "placements_enumerated_per_s": placements / seconds if placements else None}


def parity_gate(w, sample_modules, device):
    """before anything is timed, the GPU search of the CPU sample must give the very insertion
    sites the reference's own matcher gives (multiset of sites; the reference's order is a thread race). A mismatch
    aborts the benchmark: a number for a wrong answer is not a number."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import util
    import biorseo_b200 as b
    seq, mask = w["seqs"][0], w["masks"][0]
    with tempfile.TemporaryDirectory() as tmp:
        lib, what = write_sample(w, tmp, sample_modules)
        try:
            expect = util.best_oracle(w["source"], lib, seq, mask, timeout=900, threads=max(2, (os.cpu_count() or 2) - 1))
        except RuntimeError as exc:
            return {"checked_modules": 0, "equal": None, "error": str(exc)[:200]}
        host = b.HostLibrary.load(w["source"], lib)
        dev = host.to_device(device)
        sites = dev.search(seq, mask)
        mine = sorted(host.records_text(sites.records()))
        placements = dev.placements_enumerated()
        res = {"checked_modules": host.n_modules, "sample": what + f", query 0 ({len(seq)} nt)", "sites": len(mine), "equal": mine == expect,
               "oracle": "reference (oracle/_ref/biorseo_ref: the reference's unmodified sources)" if util.have_ref() else "port (oracle/_build/biorseo_port)",
               "gpu_placements_enumerated": placements}
        sites.close()
        dev.close()
    if not res["equal"]:
        sys.stderr.write(f"bench: PARITY FAILURE on {w['name']}: {len(mine)} GPU sites vs {len(expect)} reference sites\n")
        print(json.dumps({"metric": METRIC, "value": None, "parity": res, "error": "GPU insertion sites differ from the reference's"}))
        sys.exit(3)
    return res


def sample_size(name):
    # bounded samples: about 1.5 s of Pool-threaded CPU search per pass on a 16-core host
    return {"c4": 15000, "c4_400k": 15000, "c4a": 50000, "c4b": 8000, "c4b_all": 800, "c3": 1000, "c2": 400, "c2_narrow": 400, "c1": 1000, "c4_desc": 500}[name]


# ----------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c4")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--scaling", default=None, choices=["strong", "weak"],
                    help="N > 1: strong (default for the JSON workloads) = one library cut over the GPUs; weak = one library per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads reported under 'also'")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    n_gpus = world if args.impl == "b200" else args.gpus
    scaling = args.scaling or ("strong" if n_gpus > 1 and args.workload in ("c4", "c4a", "c4b", "c4b_all", "c4_400k", "c3", "c1") else "weak")
    if n_gpus == 1:
        scaling = "weak"     # (one GPU: the two coincide; the contract's default label)

    def config_of(w, units_total):
        """The same keys and values in both arms (the driver compares them); the reference arm adds its sample."""
        return {"workload": args.workload, "describe": w["describe"], "units_total": int(units_total), "scaling": scaling,
                "parallelism": (f"library-sharded x{n_gpus}, {'one library cut by search cost' if scaling == 'strong' else 'one library per GPU'}, "
                                "all-gather-v of site records") if n_gpus > 1 else "single GPU",
                "l2": "flushed between steps (256 MiB write, outside the timed events)",
                "instrumentation": "the library's per-phase CUDA events are off in the timed regions; per-kernel times come from extra untimed steps with them on"}

    if args.impl == "reference":
        if rank != 0:
            return
        w = make_workload(args.workload, 0)
        res = cpu_search(w, sample_size(args.workload), repeat=max(1, min(args.steps, 20)))   # about 30 s of CPU search at most
        if res is None:
            print(json.dumps({"impl": "reference", "unavailable": "neither oracle/_ref/biorseo_ref nor oracle/_build/biorseo_port is built"}))
            return
        import biorseo_b200 as b
        with tempfile.TemporaryDirectory() as tmp:     # (the host loaders run without a GPU: the library's unit count for the config keys)
            units_total = int(host_library(w, tmp)[0].table()["n_units"])
        if scaling == "weak" and n_gpus > 1:
            units_total *= n_gpus
        line = {"metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * res["seconds_per_pass"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
                "dtype": "u32", "data": "synthetic", "impl": "reference",
                "config": {**config_of(w, units_total), "sample_modules": res["sample_modules"], "reference_threads": res["cores"]},
                "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "sites_per_s": res["sites_per_s"], "placements_enumerated_per_s": res["placements_enumerated_per_s"], "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    import biorseo_b200 as b
    from biorseo_b200 import sharded

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    bench_stream = torch.cuda.Stream()
    torch.cuda.set_stream(bench_stream)

    exchange_name = {"value": None}

    def run_workload(name, steps, warmup, with_e2e=True, with_collective=True, scaling="weak"):
        """scaling (world > 1): "weak" = every rank searches its own library of the named size; "strong" = ONE library
        (rank 0's), cut into contiguous module ranges balanced by the search-cost estimate, one range per rank."""
        strong = scaling == "strong" and world > 1
        w = make_workload(name, 0 if strong else rank)
        if strong and w["source"] != "jsonfolder":
            raise SystemExit("strong scaling is implemented for the JSON workloads")
        with tempfile.TemporaryDirectory() as tmp:
            host, _ = host_library(w, tmp)
        table = host.table()
        seq_lens = [len(s) for s in w["seqs"]]
        units_total, shard = int(table["n_units"]), (0, host.n_modules)
        if strong:
            costs = sharded.module_costs(table, max(seq_lens))
            shard = sharded.shard_ranges_by_cost(costs, world)[rank]
            host = b.HostLibrary.from_json_modules(w["modules"][shard[0]:shard[1]])
            table = host.table()
        dev = host.to_device(local_rank)
        dev.set_module_base(shard[0] if strong else rank * host.n_modules)
        stream = bench_stream                # a real (non-default) stream: kernels, collectives and the timing events all go there
        dev.set_stream(stream.cuda_stream)
        query = dev.query(w["seqs"], w["masks"])
        tested_per_step = sum(seq_lens) * int(table["n_units"])        # this rank's share of a step
        tested_total = sum(seq_lens) * units_total if strong else None  # the whole job of a step (strong: fixed as N grows)

        # size the output once (capacity negotiation of the C ABI), then reuse it every step
        rc, n_words, n_sites = dev.search_query_into(query, 0, 0)
        records = torch.empty(max(n_words, 1), dtype=torch.int32, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

        # Library shards of one box hold similar numbers of sites: the search writes header + records
        # straight into this rank's block of a padded single-collective exchange. Very large streams
        # take the exact-size point-to-point all-gather-v instead (no padding, no final cut).
        # (capacity: the largest shard stream of the job with 25 % head-room; the peer-memory exchange holds 2 x world x capacity words)
        cap_words = int(1.25 * n_words) + 1024
        if world > 1:
            t = torch.tensor([cap_words], dtype=torch.int64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cap_words = int(t.item())
        padded = world > 1 and with_collective and 8 * world * cap_words <= (48 << 30)
        use_peer = padded and os.environ.get("BSS_EXCHANGE", "p2p") == "p2p"
        gatherer = sharded.PaddedGatherer(torch.device("cuda", local_rank), n_seqs=len(w["seqs"]), capacity_words=cap_words) if padded and not use_peer else None
        # The exchange of choice: ONE kernel over NVLink peer memory (counts, records and completion flags as peer-to-peer
        # stores); the padded NCCL all-gather + cut kernel when peer mapping is unavailable (BSS_EXCHANGE=nccl forces it).
        peer = None
        if use_peer:
            try:
                peer = sharded.PeerGatherer(torch.device("cuda", local_rank), n_seqs=len(w["seqs"]), capacity_words=cap_words)
            except RuntimeError as exc:
                if rank == 0:
                    sys.stderr.write(f"bench: {exc}; using the NCCL all-gather\n")
                gatherer = sharded.PaddedGatherer(torch.device("cuda", local_rank), n_seqs=len(w["seqs"]), capacity_words=cap_words)
        exchange_name["value"] = "peer-memory kernel (bss_exchange)" if peer is not None else "NCCL all_gather_into_tensor + cut kernel" if padded else "NCCL grouped send/recv"

        gathered = []
        breakdown = []   # one event per step between the search and the exchange: per-rank search / exchange times

        def step():
            """Queues one whole step on the stream — search (count -> scan -> emit as one CUDA graph) and, for
            N > 1, the exchange — without waiting for the host anywhere; finish() then checks what it produced."""
            if peer is not None:
                head_ptr, rec_ptr, cap = peer.block_pointers()
                dev.search_query_into_async(query, rec_ptr, cap, seq_begin_ptr=head_ptr)
                breakdown.append(torch.cuda.Event(enable_timing=True))
                breakdown[-1].record(stream)
                gathered.append(peer.exchange_async(stream.cuda_stream))
                if len(gathered) > 2:
                    gathered.pop(0)     # the output is double-buffered: only the last two are valid
                return
            if padded:
                # search straight into this rank's block, all-gather, cut
                # (the block was sized from the first search; an overflow would show in the flag checked after the loop)
                head, rec = gatherer.block()
                dev.search_query_into_async(query, rec.data_ptr(), rec.numel(), seq_begin_ptr=head.data_ptr())
                breakdown.append(torch.cuda.Event(enable_timing=True))
                breakdown[-1].record(stream)
                gathered.append(gatherer.exchange_async(dev, stream.cuda_stream))
                return
            dev.search_query_into_async(query, records.data_ptr(), records.numel())
            breakdown.append(torch.cuda.Event(enable_timing=True))
            breakdown[-1].record(stream)
            if world > 1 and with_collective:
                sharded.allgatherv(records[:n_words])

        def finish():
            rc, nw, ns = dev.search_wait()
            assert rc == 0 and nw == n_words and ns == n_sites
            return dev.stats()

        for _ in range(warmup):
            step()
            finish()
        if peer is not None:
            # The peer-memory exchange proves itself on the warm-up steps before it is timed: no flag raised and the very
            # concatenation NCCL's exact-size all-gather-v gives. Otherwise every rank drops to the NCCL exchange together.
            out, result = gathered[-1]
            result = result.tolist()
            ok = result[world + 1] == 0 and result[rank] == n_words
            want, counts = sharded.allgatherv(peer.records(n_words).clone())
            ok = ok and counts == result[:world] and torch.equal(out[:sum(counts)], want)
            agree = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
            dist.all_reduce(agree, op=dist.ReduceOp.MIN)
            if not int(agree.item()):
                if rank == 0:
                    sys.stderr.write("bench: the peer-memory exchange failed its warm-up check; using the NCCL all-gather\n")
                peer.close()
                peer = None
                gatherer = sharded.PaddedGatherer(torch.device("cuda", local_rank), n_seqs=len(w["seqs"]), capacity_words=cap_words)
                exchange_name["value"] = "NCCL all_gather_into_tensor + cut kernel (peer-memory exchange failed its check)"
                gathered.clear()
                for _ in range(warmup):
                    step()
                    finish()
        # The library's CUDA events BETWEEN the phases of a search (count / scan / emit times of bss_stats) are instrumentation: off in
        # the timed regions (they cost ~7 us of a 0.16 ms step as extra graph nodes), on again for the per-kernel times taken afterwards.
        dev.set_tuning("phase_events", 0)
        for _ in range(2):
            step()
            finish()
        sampler = ClockSampler(local_rank)
        sampler.start()
        kernel_ms = {"count_ms": [], "scan_ms": [], "emit_ms": []}
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall0 = time.perf_counter()
        # The K timed steps are queued back to back — flush, event, search (+ exchange), event — and the host waits once,
        # behind the last one: the device never idles on the host and, for N > 1, the ranks stay in step with each other
        # through their exchanges instead of drifting apart by host jitter. Each step is timed by its own pair of events
        # (the flush in between is outside them).
        events = []
        breakdown.clear()
        host_t0 = time.perf_counter()
        for _ in range(steps):
            ef = torch.cuda.Event(enable_timing=True)
            ef.record(stream)
            flush.fill_(1)                      # evict the previous step's lines from L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step()
            e1.record(stream)
            events.append((e0, breakdown[-1], e1, ef))
        host_enqueue_ms = 1e3 * (time.perf_counter() - host_t0) / steps     # host time to queue one step (the device runs behind)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        wall = time.perf_counter() - wall0
        st = finish()                           # sizes of the last search checked (every step runs the same search)
        step_ms = [a.elapsed_time(z) for a, _, z, _ in events]
        search_ms = float(np.mean([a.elapsed_time(m) for a, m, _, _ in events]))      # this rank's search, device time
        exchange_ms = float(np.mean([m.elapsed_time(z) for _, m, z, _ in events]))    # exchange kernel, waiting for the slowest peer included
        flush_ms = float(np.mean([f.elapsed_time(a) for a, _, _, f in events]))       # the L2 flush between steps (outside the step's events)
        # device time from the end of a step to the start of the next one's flush: > 0 means the device waited for the host to queue work
        idle_ms = float(np.mean([events[i][2].elapsed_time(events[i + 1][3]) for i in range(len(events) - 1)])) if len(events) > 1 else 0.0
        launches = steps * st["kernel_launches"]
        # per-kernel device times (the library's own events belong to the last search queued): a few more steps, one at a time, untimed
        dev.set_tuning("phase_events", 1)
        for _ in range(2):
            step()
            finish()
        for _ in range(3):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            step()
            e1.record(stream)
            torch.cuda.synchronize()
            st = finish()
            for k in kernel_ms:
                kernel_ms[k].append(st[k])
        if gathered and peer is not None:   # complete and consistent: no overflow, no timeout, every rank's count, canonical order
            out, result = gathered[-1]
            result = result.tolist()
            assert result[world + 1] == 0, f"exchange flags {result[world + 1]}"
            counts = result[:world]
            assert counts[rank] == n_words and result[world] == sum(counts)
            mine = out[sum(counts[:rank]):sum(counts[:rank + 1])]
            assert torch.equal(mine, peer.records(n_words))
            every = [torch.empty(1, dtype=torch.int64, device="cuda") for _ in range(world)]
            dist.all_gather(every, out[:sum(counts)].to(torch.int64).sum().reshape(1))     # every rank holds the same concatenation
            assert all(int(e.item()) == int(every[0].item()) for e in every)
            launches += steps    # one exchange kernel per step
        elif gathered:   # the exchanges were complete and consistent: no overflow, every rank's count, canonical order
            out, counts, overflow = gathered[-1]
            assert int(overflow.item()) == 0
            counts = counts.tolist()
            assert counts[rank] == n_words
            mine = out[sum(counts[:rank]):sum(counts[:rank + 1])]
            assert torch.equal(mine, gatherer.block()[1][:n_words])
        clocks = sampler.summary()

        total_ms = float(sum(step_ms))
        per_rank, sites_all = None, n_sites * world
        if world > 1:
            t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total_ms = float(t.item())
            # per-rank breakdown: units, sites, search time, exchange time (its wait for the slowest peer included), and the
            # skew = how much earlier than the slowest rank this rank's search ended
            mine = torch.tensor([int(table["n_units"]), n_sites, n_words, search_ms, exchange_ms, float(sum(step_ms)) / steps, flush_ms, idle_ms, host_enqueue_ms],
                                dtype=torch.float64, device="cuda")
            every = torch.empty(world * mine.numel(), dtype=torch.float64, device="cuda")
            dist.all_gather_into_tensor(every, mine)
            rows = every.view(world, -1).tolist()
            slowest = max(r[3] for r in rows)
            words_all = sum(r[2] for r in rows)
            per_rank = [{"rank": i, "units": int(r[0]), "sites": int(r[1]), "words": int(r[2]), "search_ms": r[3], "exchange_ms": r[4],
                         "skew_ms": slowest - r[3], "step_ms": r[5], "flush_ms": r[6], "device_idle_between_steps_ms": r[7], "host_enqueue_ms": r[8],
                         # what the rank receives over the links in one exchange / its exchange time (the wait for the slowest peer included)
                         "exchange_in_gbs": 4.0 * (words_all - r[2]) / (r[4] * 1e-3) / 1e9 if r[4] > 0 else None} for i, r in enumerate(rows)]
            sites_all = int(sum(r[1] for r in rows))
        job_tested = tested_total if strong else world * tested_per_step
        link = None
        if world > 1 and with_collective and 4 * (words_all if per_rank else 0) >= (64 << 20):
            # the box's own peer-copy rate, measured here: every rank sends 256 MiB to its right neighbour and receives from its left one (NCCL)
            buf_out, buf_in = torch.empty(64 << 20, dtype=torch.int32, device="cuda"), torch.empty(64 << 20, dtype=torch.int32, device="cuda")
            ms = []
            for i in range(4):
                a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for req in dist.batch_isend_irecv([dist.P2POp(dist.isend, buf_out, (rank + 1) % world), dist.P2POp(dist.irecv, buf_in, (rank - 1) % world)]):
                    req.wait()
                z.record(stream)
                torch.cuda.synchronize()
                if i:
                    ms.append(a.elapsed_time(z))
            link = {"measured_peer_copy_gbs_per_direction": (256 << 20) / (min(ms) * 1e-3) / 1e9, "how": "NCCL send/recv ring of 256 MiB per rank, best of 3"}
            del buf_out, buf_in
        res = {"workload": w, "table": table, "value": job_tested * steps / (total_ms * 1e-3), "scaling": "strong" if strong else "weak", "link": link,
               "ms_per_step": total_ms / steps, "sites": n_sites, "words": n_words, "launches": launches, "clocks": clocks,
               "kernel_ms": {k: float(np.mean(v)) for k, v in kernel_ms.items()}, "wall_s": wall, "tested_per_step": tested_per_step,
               "sites_per_s": sites_all * steps / (total_ms * 1e-3), "seq_lens": seq_lens, "per_rank": per_rank, "units_total": units_total,
               "stats": {k: st[k] for k in ("window_items", "ordinary_items", "heavy_items")},
               "limiter": None}
        if per_rank:
            worst = max(per_rank, key=lambda r: r["search_ms"])
            res["limiter"] = (f"slowest search {worst['search_ms']:.3f} ms on rank {worst['rank']} ({worst['units']} units), mean exchange "
                              f"{np.mean([r['exchange_ms'] for r in per_rank]):.3f} ms of which skew {np.mean([r['skew_ms'] for r in per_rank]):.3f} ms")

        try:     # the reference's real work term for the whole workload (a diagnostic kernel of its own, outside the timed region)
            res["placements_enumerated"] = int(dev.placements_enumerated(query))
            res["placements_enumerated_per_s"] = world * res["placements_enumerated"] * steps / (total_ms * 1e-3) if not strong else None
        except Exception as exc:
            res["placements_enumerated"] = f"unavailable: {exc}"
        if with_e2e:
            dev.set_tuning("phase_events", 0)
            # end to end through the host-buffer C ABI: pinned host inputs -> device -> pinned host records
            batch = b.HostBatch(w["seqs"], w["masks"], pinned=True)   # packed once, page-locked: the C ABI's input form
            for _ in range(2):
                dev.search_host(batch).close()
            e2e_steps = max(3, min(steps, 10))
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                s = dev.search_host(batch)     # H2D of sequences + masks, kernels, D2H of the records: all inside
                assert s.n_words == n_words
                s.close()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            st = dev.stats()
            res["e2e"] = {"value": job_tested * e2e_steps / dt, "unit": UNIT, "h2d_bytes_per_step": int(query.h2d_bytes),
                          "d2h_bytes_per_step": int(st["d2h_bytes"]), "ms_per_step": 1e3 * dt / e2e_steps,
                          "path": "bss_search_batch (C ABI): pinned host sequences + masks in, pinned host records out"}
        query.close()
        dev.close()
        if peer is not None:
            peer.close()
        return res

    # parity gate: the GPU search of the CPU sample against the reference's own matcher, before timing
    parity = None
    if rank == 0:
        parity = parity_gate(make_workload(args.workload, 0), sample_size(args.workload), local_rank)
    if world > 1:
        dist.barrier()

    main_res = run_workload(args.workload, args.steps, args.warmup, scaling=scaling)
    w, table = main_res["workload"], main_res["table"]
    weak_res = None
    if world > 1 and scaling == "strong":      # the weak-scaling number beside it (one library of the named size per GPU)
        weak_res = run_workload(args.workload, max(3, args.steps // 2), args.warmup, with_e2e=False, scaling="weak")

    # roofline of the dominant kernel (HBM: this path has no contraction, tensor cores are not used)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    traffic_table = {}
    for name in ("r01_traffic.json", "r02_traffic.json"):    # DRAM bytes per launch from the committed ncu captures; the later round wins
        path = os.path.join(ROOT, "profiles", name)
        if os.path.exists(path):
            for wl, per_kernel in json.load(open(path)).items():
                if isinstance(per_kernel, dict):
                    traffic_table.setdefault(wl, {}).update(per_kernel)

    def roofline(res):
        km = res["kernel_ms"]
        name = max(km, key=km.get)
        traffic = traffic_table.get(res["workload"]["name"], {}).get(name)   # DRAM bytes per launch, from the committed ncu capture
        b_in, b_out = algorithmic_bytes(res["table"], res["seq_lens"], res["words"])
        # the count kernel reads the inputs, the emit kernel reads them again and writes the sites
        algo = {"count_ms": b_in, "scan_ms": 0, "emit_ms": b_in + b_out}[name]
        achieved = algo / (km[name] * 1e-3) / 1e9 if km[name] > 0 else 0.0
        windowed = res["stats"]["window_items"] > 0
        return {"bound": "hbm", "kernel": {"count_ms": "pattern_kernel + route_kernel + window_kernel<count> (pattern table, routing, window walk)" if windowed
                                           else "search_kernel<G,false> (match+count)", "scan_ms": "scan_kernel",
                                           "emit_ms": "window_kernel<emit>" if windowed else "search_kernel<G,true> (emit)"}[name],
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "algorithmic_bytes_per_launch": algo, "kernel_ms": km[name], "peak_source": peak_src,
                "all_kernels_ms": km}

    int_peak = {"value": None}

    def roofline_int(res):
        """Second roofline of the search: thread-level shift / logic operations against the rate an SHF + LOP3 only
        kernel sustains on this GPU (measured here, bss_measure_int_issue). `achieved` counts ALGORITHMIC operations
        (the shift-AND lower bound), so the fraction says how far the whole pass is from a pure bit-parallel scan;
        the executed instruction counts are in profiles/ (ncu)."""
        if int_peak["value"] is None:
            try:
                host = b.HostLibrary.from_json_modules(make_workload("c1", 0)["modules"][:8])
                dev = host.to_device(local_rank)
                int_peak["value"] = dev.measure_int_issue()
                dev.close()
            except Exception:
                int_peak["value"] = 0.0
        km = res["kernel_ms"]
        ops = algorithmic_int_ops(res["table"], res["seq_lens"], res["sites"])
        t = (km["count_ms"] + km["emit_ms"]) * 1e-3
        achieved = ops / t / 1e9 if t > 0 else 0.0
        peak_g = int_peak["value"] / 1e9
        return {"bound": "int_issue", "kernels": "prefilter + count + emit", "achieved": achieved, "peak": peak_g, "unit": "Gop/s (thread-level SHF/LOP3)",
                "frac": achieved / peak_g if peak_g else None, "algorithmic_ops_per_step": ops, "kernel_ms": km["count_ms"] + km["emit_ms"],
                "peak_source": "bss_measure_int_issue: SHF + LOP3 only, 8 independent chains per thread, 8 CTAs x 256 threads per SM, measured in this run"}

    units_total = main_res["units_total"] * (world if scaling == "weak" and world > 1 else 1)
    line = {"metric": METRIC, "value": main_res["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "u32", "data": "synthetic",
            "config": config_of(w, units_total),
            "exchange": exchange_name["value"] if world > 1 else None, "parity": parity,
            "per_rank": main_res["per_rank"], "limiter": main_res["limiter"], "link": main_res.get("link"), "routing": main_res["stats"],
            "weak_scaling": None if weak_res is None else {"value": weak_res["value"], "ms_per_step": weak_res["ms_per_step"], "units_per_gpu": weak_res["units_total"],
                                                           "sites_per_s": weak_res["sites_per_s"], "per_rank": weak_res["per_rank"], "limiter": weak_res["limiter"]},
            "sites_per_s": main_res["sites_per_s"], "sites_per_step_per_gpu": main_res["sites"],
            "placements_enumerated_per_step_per_gpu": main_res.get("placements_enumerated"),
            "roofline": roofline(main_res), "roofline_int": roofline_int(main_res) if rank == 0 else None, "clocks": main_res["clocks"], "e2e": main_res.get("e2e"),
            "gpu_launches": main_res["launches"], "wall_s_timed_region": main_res["wall_s"]}

    if rank == 0 and world == 1 and not args.no_extra:
        also = {}
        for name in ("c4b_all", "c3", "c2", "c2_narrow", "c1", "c4_desc"):
            if name == args.workload:
                continue
            try:
                r = run_workload(name, max(3, args.steps // 4), 3, with_e2e=True)
                also[name] = {"describe": r["workload"]["describe"], "value": r["value"], "ms_per_step": r["ms_per_step"], "routing": r["stats"],
                              "placements_enumerated": r.get("placements_enumerated"), "parity": parity_gate(r["workload"], min(sample_size(name), 200), local_rank),
                              "sites_per_s": r["sites_per_s"], "sites_per_step": r["sites"], "e2e": r.get("e2e"), "roofline": roofline(r),
                              "roofline_int": roofline_int(r)}
            except Exception as exc:   # a secondary workload must never take the headline down
                also[name] = {"error": str(exc)[:200]}
        line["also"] = also

    if rank == 0 and world == 1 and not args.no_extra:
        # Pair-mask producer (row "next" of the path, cppsrc/MOIP.cpp:77-90): n^2 floats in, n^2/8 bytes out,
        # device pointers, CUDA events; 16000 nt (BSS_MAX_SEQ_LEN): the 1 GB matrix is far larger than L2.
        try:
            host = b.HostLibrary.from_json_modules(make_workload("c1", 0)["modules"][:8])
            dev = host.to_device(local_rank)
            n = 16000
            pij = torch.rand((n, n), dtype=torch.float32, device="cuda") * 0.004
            out = torch.empty((n, (n + 31) // 32), dtype=torch.int32, device="cuda")
            stream = bench_stream
            res = {}
            for layout, (rs, cs) in (("row_major", (n, 1)), ("column_major", (1, n))):
                for _ in range(3):
                    dev.pair_mask_device(pij.data_ptr(), n, rs, cs, 1e-3, out.data_ptr(), stream.cuda_stream)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(10):
                    dev.pair_mask_device(pij.data_ptr(), n, rs, cs, 1e-3, out.data_ptr(), stream.cuda_stream)
                e1.record(stream)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                nbytes = (n * (n - 1) // 2) * 4 + n * ((n + 31) // 32) * 4      # the floats above the diagonal + the whole bit matrix
                res[layout] = {"ms": ms, "achieved": nbytes / (ms * 1e-3) / 1e9, "unit": "GB/s", "frac": nbytes / (ms * 1e-3) / 1e9 / peak,
                               "algorithmic_bytes_per_launch": nbytes}
            line["also"]["pair_mask_from_pij_16000"] = {"describe": "allowed_basepair bit matrix from a 16000 x 16000 float32 probability matrix (HBM-bound: the floats above the diagonal in, the bit matrix out)",
                                                        "bound": "hbm", "peak": peak, **res}
            dev.close()
        except Exception as exc:
            line.setdefault("also", {})["pair_mask_from_pij_16000"] = {"error": str(exc)[:200]}

    if rank == 0 and world == 1 and not args.no_extra:
        # Insertion-variable index (row "next" #1: cppsrc/MOIP.cpp:302-328 variables, :507-586 c3 / c4 from the site
        # list), built on the device right behind a search. HBM/L2-bound integer streams: records + pair mask in,
        # variables table + overlap lists + pair adjacency out. Device time by CUDA events inside the library.
        def index_bench(name, modules, mask):
            seq = synth_mod.sequence(2000, 4)
            host = b.HostLibrary.from_json_modules(modules)
            dev = host.to_device(local_rank)
            dev.set_stream(bench_stream.cuda_stream)
            q = dev.query([seq], [mask])
            sites = dev.search_query(q, fetch=False)
            ms, launches, idx = [], 0, None
            for i in range(6):
                flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda").fill_(1)
                del flush
                idx = dev.site_index(q, sites.device_pointer(), 0, fetch=False)
                if i >= 2:
                    ms.append(idx.build_ms)
                    launches += idx.kernel_launches
                counts = {sec: idx.count(sec) for sec in b.capi.IDX_SECTIONS}
                idx.close()
            n = len(seq)
            nbytes = 4 * sites.n_words + n * ((n + 31) // 32) * 4 + 4 * sum(counts.values())
            t = float(np.mean(ms))
            res = {"describe": name, "sites": sites.n_sites, "variables": counts["var_first"], "overlap_entries": counts["overlap_vars"],
                   "allowed_pairs": counts["pair_partner"] // 2, "ms": t, "sites_per_s": sites.n_sites / (t * 1e-3), "bound": "hbm",
                   "algorithmic_bytes": nbytes, "achieved": nbytes / (t * 1e-3) / 1e9, "unit": "GB/s", "peak": peak,
                   "frac": nbytes / (t * 1e-3) / 1e9 / peak, "gpu_launches": launches,
                   "note": "device timeline of the whole build, two host round trips for the sizes included"}
            sites.close()
            q.close()
            dev.close()
            return res
        def index_cpu(modules, mask):
            """The reference's own host loops for the same thing (cppsrc/MOIP.cpp:307-328, 507-586), restated without the
            solver objects in oracle/port (`index` mode) and timed on one host core on a bounded sample of the library."""
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import util
            if not util.have_port():
                return None
            with tempfile.TemporaryDirectory() as tmp:
                lib = synth_mod.write_json(os.path.join(tmp, "lib.json"), modules)
                seq_path, mask_path = util.write_query(tmp, synth_mod.sequence(2000, 4), mask)
                out = subprocess.run([util.PORT_BIN, "index", "jsonfolder", lib, seq_path, mask_path], capture_output=True, text=True, check=True).stdout
            r = json.loads(out.strip().splitlines()[-1])
            return {"kind": "port", "cores": 1, "sites": r["sites"], "seconds": r["seconds"], "sites_per_s": r["sites"] / r["seconds"],
                    "sample": f"{len(modules)} modules of the same library, host loops only (search excluded)"}
        try:
            from biorseo_b200 import synth as synth_mod
            seq4 = synth_mod.sequence(2000, 4)
            line["also"]["site_index"] = {
                "c4": index_bench("index of the c4 sites (2000 nt x 50000 units, banded mask)", make_workload("c4", 0)["modules"],
                                  synth_mod.mask_canonical(seq4, seed=204, keep=0.3, span=100)),
                "c4b_all_2000": index_bench("index of 2000 units of c4b with the all-allowed mask (output-heavy)",
                                            synth_mod.json_modules(2000, 104, "c4b"), synth_mod.mask_all(2000)),
            }
            if not args.no_cpu_baseline:
                line["also"]["site_index"]["c4b_all_2000"]["cpu_baseline"] = index_cpu(synth_mod.json_modules(2000, 104, "c4b")[:30], synth_mod.mask_all(2000))
        except Exception as exc:
            line.setdefault("also", {})["site_index"] = {"error": str(exc)[:300]}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cb = cpu_search(w, sample_size(args.workload), repeat=3)
        line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "host_cores", "kind", "sample", "sites_per_s", "placements_enumerated",
                                                   "placements_enumerated_per_s")} if cb else None
        # the same search on ONE thread (SURVEY 8d asks for both), on a smaller sample of the library
        c1t = cpu_search(w, max(1, sample_size(args.workload) // 10), repeat=1, threads=1)
        line["cpu_baseline_single_thread"] = {k: c1t[k] for k in ("value", "unit", "cores", "host_cores", "kind", "sample", "sites_per_s")} if c1t else None

    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
