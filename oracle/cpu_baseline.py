"""CPU arm of bench.py, run as a SEPARATE PROCESS with pinned OpenMP threads (test infrastructure, like the rest of oracle/).

    python oracle/cpu_baseline.py DATASET.npy QUERIES.npy K REPEATS [scan|blocked|auto]

Times exact fp32 kNN of the given queries over the given rows with the oracle port (sequential-fmaf scan `oracle.knn`, or
the blocked SGEMM + top-k formulation `oracle.knn_blocked`, whichever is faster on this machine unless forced), REPEATS times
after one warm-up, and prints one JSON object: {"times_s": [...], "formulation": ..., "threads": ..., "queries": ..., "rows": ...}.
bench.py launches it with OMP_NUM_THREADS / OMP_PROC_BIND / OMP_PLACES set and nothing else running in the process: the
round-1 in-process measurement varied 5x between runs (thread pool shared with the GPU arm's host threads).
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np  # noqa: E402

import oracle  # noqa: E402


def main():
    ds = np.load(sys.argv[1], mmap_mode="r")
    qs = np.load(sys.argv[2])
    k, reps = int(sys.argv[3]), int(sys.argv[4])
    mode = sys.argv[5] if len(sys.argv) > 5 else "auto"
    threads = int(os.environ.get("OMP_NUM_THREADS", os.cpu_count() or 1))
    oracle.set_threads(threads)
    ds = np.ascontiguousarray(ds)
    probe = min(16, len(qs))
    rates = {}
    if mode in ("auto", "scan"):
        t0 = time.time()
        oracle.knn(ds, qs[:4], k)
        rates["scan"] = 4 / max(time.time() - t0, 1e-9)
    if mode in ("auto", "blocked"):
        oracle.knn_blocked(ds, qs[:probe], k, threads=threads)  # warm-up (thread pool, page faults)
        t0 = time.time()
        oracle.knn_blocked(ds, qs[:probe], k, threads=threads)
        rates["blocked"] = probe / max(time.time() - t0, 1e-9)
    form = max(rates, key=rates.get)
    run = (lambda: oracle.knn_blocked(ds, qs, k, threads=threads)) if form == "blocked" else (lambda: oracle.knn(ds, qs, k))
    run()
    times = []
    for _ in range(reps):
        t0 = time.time()
        run()
        times.append(time.time() - t0)
    print(json.dumps({"times_s": times, "formulation": form, "probe_rates_qps": rates, "threads": threads,
                      "queries": int(len(qs)), "rows": int(ds.shape[0])}))


if __name__ == "__main__":
    main()
