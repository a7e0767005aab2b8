"""cuvs-bench backend plugin: runs the reference's benchmark harness on this library, in process.

Mirrors the plugin interface of the reference's benchmark package
  python/cuvs_bench/cuvs_bench/backends/base.py:17-338   Dataset / BuildResult / SearchResult / BenchmarkBackend
  python/cuvs_bench/cuvs_bench/backends/registry.py      get_registry().register(name, cls)
  python/cuvs_bench/cuvs_bench/orchestrator/config_loaders.py:26-49   IndexConfig(name, algo, build_param, search_params, file)
and the parameter names of the reference's C++ wrappers `algo<T>` (cpp/bench/ann/src/common/ann_types.hpp:124-166,
cpp/bench/ann/src/cuvs/cuvs_ann_bench_param_parser.h: nlist / niter / ratio / pq_dim / pq_bits / nprobe / refine_ratio /
graph_degree / intermediate_graph_degree / itopk / search_width / internalDistanceDtype / smemLutDtype), so the YAML configs
under python/cuvs_bench/cuvs_bench/config/algos/cuvs_*.yaml drive it unchanged:

    from cuvs_bench.backends import get_registry
    import cuvs_b200.bench_backend as bb
    bb.register()                                    # name "cuvs_b200"
    backend = get_registry().get_backend("cuvs_b200", {"name": "cuvs_ivf_pq.nlist1024"})
    backend.build(dataset, indexes); backend.search(dataset, indexes, k=10, batch_size=10000, mode="throughput")

When the reference's `cuvs_bench` package is importable its base classes are used (so `isinstance` checks of the orchestrator
hold); otherwise structurally identical local definitions stand in (the GPU box has no reference checkout).  Algorithms:
cuvs_brute_force, cuvs_ivf_flat, cuvs_ivf_pq, cuvs_cagra — GPU only, served by libcuvs_c.so through the ctypes binding.
There is no CPU algorithm here: BASELINE configs[0] (hnswlib through cuvs_bench, CPU, "plumbing") is covered by
tests/test_bench_backend_cpu.py, which drives the SAME harness code with a test-only CPU backend.
"""
from __future__ import annotations

import os
import time
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional

import numpy as np

try:  # the reference's own plugin base classes, when its benchmark package is on the path
    from cuvs_bench.backends.base import BenchmarkBackend, BuildResult, Dataset, SearchResult  # type: ignore
    HAVE_CUVS_BENCH = True
except Exception:  # noqa: BLE001 - absent package, or one of its optional imports
    HAVE_CUVS_BENCH = False

    class Dataset:  # python/cuvs_bench/cuvs_bench/backends/base.py:17-201 (array-based usage pattern)
        def __init__(self, name, training_vectors=None, query_vectors=None, groundtruth_neighbors=None,
                     groundtruth_distances=None, distance_metric="euclidean", base_file=None, query_file=None,
                     groundtruth_neighbors_file=None, groundtruth_distances_file=None, metadata=None):
            self.name = name
            self.training_vectors = training_vectors if training_vectors is not None else np.empty((0, 0))
            self.query_vectors = query_vectors if query_vectors is not None else np.empty((0, 0))
            self.groundtruth_neighbors = groundtruth_neighbors
            self.groundtruth_distances = groundtruth_distances
            self.distance_metric = distance_metric
            self.base_file, self.query_file = base_file, query_file
            self.groundtruth_neighbors_file, self.groundtruth_distances_file = groundtruth_neighbors_file, groundtruth_distances_file
            self.metadata = metadata or {}

        @property
        def dims(self):
            return 0 if self.training_vectors.size == 0 else self.training_vectors.shape[1]

        @property
        def n_base(self):
            return self.training_vectors.shape[0]

        @property
        def n_queries(self):
            return self.query_vectors.shape[0]

    @dataclass
    class BuildResult:  # base.py:203-257
        index_path: str
        build_time_seconds: float
        index_size_bytes: int
        algorithm: str
        build_params: Dict[str, Any]
        metadata: Dict[str, Any] = field(default_factory=dict)
        success: bool = True
        error_message: Optional[str] = None

        def to_json(self):
            return {"name": f"{self.algorithm}/build", "real_time": self.build_time_seconds, "time_unit": "s",
                    "index_size": self.index_size_bytes, "success": self.success, **self.build_params, **self.metadata}

    @dataclass
    class SearchResult:  # base.py:260-337
        neighbors: np.ndarray
        distances: np.ndarray
        search_time_ms: float
        queries_per_second: float
        recall: float
        algorithm: str
        search_params: List[Dict[str, Any]]
        latency_percentiles: Optional[Dict[str, float]] = None
        gpu_time_seconds: Optional[float] = None
        cpu_time_seconds: Optional[float] = None
        metadata: Dict[str, Any] = field(default_factory=dict)
        success: bool = True
        error_message: Optional[str] = None

        def to_json(self):
            out = {"name": f"{self.algorithm}/search", "real_time": self.search_time_ms, "time_unit": "ms",
                   "items_per_second": self.queries_per_second, "Recall": self.recall, "success": self.success,
                   "search_params": self.search_params, **self.metadata}
            if self.latency_percentiles:
                out.update(self.latency_percentiles)
            if self.gpu_time_seconds is not None:
                out["GPU"] = self.gpu_time_seconds
            if self.cpu_time_seconds is not None:
                out["cpu_time"] = self.cpu_time_seconds
            return out

    class BenchmarkBackend:  # base.py:340-520
        def __init__(self, config):
            self.config = config

        def initialize(self):
            pass

        def cleanup(self):
            pass


@dataclass
class IndexConfig:
    """Same fields as orchestrator/config_loaders.py:26-49 (any object with these attributes is accepted)."""
    name: str
    algo: str
    build_param: Dict[str, Any]
    search_params: List[Dict[str, Any]]
    file: str = ""


METRICS = {"euclidean": "sqeuclidean", "sqeuclidean": "sqeuclidean", "inner_product": "inner_product", "cosine": "cosine"}
_DTYPES = {"float": np.float32, "fp32": np.float32, "float32": np.float32, "half": np.float16, "fp16": np.float16,
           "float16": np.float16, "fp8": np.uint8, "uint8": np.uint8}


def recall_at_k(found: np.ndarray, truth: np.ndarray, k: int) -> float:
    """Set recall as the reference harness computes it (cpp/bench/ann/src/common/benchmark.hpp:300-341: hits over the first k
    ground-truth ids / (n_queries * k))."""
    found, truth = np.asarray(found)[:, :k], np.asarray(truth)[:, :k]
    hits = sum(len(np.intersect1d(f, t)) for f, t in zip(found, truth))
    return hits / float(found.shape[0] * k)


class HarnessMixin:
    """The part of the plugin that does not depend on the algorithm: batching, timing modes, recall, result records.
    Subclasses provide `_build_one(algo, metric, vectors, build_param)` -> handle, `_search_batch(handle, algo, queries, k,
    search_param)` -> (distances, neighbors) as numpy, and may override `_sync()` / `_index_bytes(handle)`."""

    requires_gpu = False
    requires_network = False

    @property
    def algo(self) -> str:  # abstract in the reference's BenchmarkBackend (base.py:561-578)
        return self.config.get("algo") or str(self.config.get("name", type(self).__name__)).split(".")[0]

    def _check_gpu_available(self) -> bool:  # the reference probes `import rmm` (base.py:484-497); this library has no RMM
        try:
            import torch
            return bool(torch.cuda.is_available())
        except ImportError:
            return False

    def _sync(self):
        pass

    def _index_bytes(self, handle) -> int:
        return 0

    def build(self, dataset, indexes, force=False, dry_run=False):
        self._handles = getattr(self, "_handles", {})
        t_total, sizes, last = 0.0, 0, None
        for ix in indexes:
            last = ix
            if dry_run:
                print(f"[dry-run] build {ix.name}: algo={ix.algo} {ix.build_param}")
                continue
            if ix.name in self._handles and not force:
                continue
            metric = METRICS[getattr(dataset, "distance_metric", "euclidean")]
            t0 = time.perf_counter()
            handle = self._build_one(ix.algo, metric, dataset.training_vectors, dict(ix.build_param))
            self._sync()
            t_total += time.perf_counter() - t0
            self._handles[ix.name] = handle
            sizes += self._index_bytes(handle)
        return BuildResult(index_path=getattr(last, "file", "") if last else "", build_time_seconds=t_total, index_size_bytes=sizes,
                           algorithm=last.algo if last else "", build_params=dict(last.build_param) if last else {},
                           metadata={"backend": self.config.get("name", type(self).__name__), "n_indexes": len(indexes)})

    def search(self, dataset, indexes, k, batch_size=10000, mode="latency", force=False, search_threads=None, dry_run=False):
        if mode not in ("latency", "throughput"):
            raise ValueError(f"mode must be 'latency' or 'throughput', got {mode!r}")
        queries = np.ascontiguousarray(dataset.query_vectors)
        nq = queries.shape[0]
        best, records = None, []
        for ix in indexes:
            if dry_run:
                print(f"[dry-run] search {ix.name}: {ix.search_params} k={k} batch={batch_size} mode={mode}")
                continue
            handle = self._handles[ix.name]
            for sp in ix.search_params or [{}]:
                neighbors = np.empty((nq, k), dtype=np.int64)
                distances = np.empty((nq, k), dtype=np.float32)
                self._search_batch(handle, ix.algo, queries[:min(nq, batch_size)], k, dict(sp), dataset)  # warm-up, as the harness's first lap
                self._sync()
                lat, t0 = [], time.perf_counter()
                for b0 in range(0, nq, batch_size):
                    tb = time.perf_counter()
                    d, i = self._search_batch(handle, ix.algo, queries[b0:b0 + batch_size], k, dict(sp), dataset)
                    if mode == "latency":
                        self._sync()
                        lat.append((time.perf_counter() - tb) * 1e3)
                    neighbors[b0:b0 + batch_size], distances[b0:b0 + batch_size] = i, d
                self._sync()
                secs = time.perf_counter() - t0
                gt = dataset.groundtruth_neighbors
                rec = recall_at_k(neighbors, gt, k) if gt is not None else float("nan")
                pct = ({"p50": float(np.percentile(lat, 50)), "p95": float(np.percentile(lat, 95)), "p99": float(np.percentile(lat, 99))}
                       if lat else None)
                res = SearchResult(neighbors=neighbors, distances=distances, search_time_ms=secs * 1e3, queries_per_second=nq / secs,
                                   recall=rec, algorithm=ix.algo, search_params=[dict(sp)], latency_percentiles=pct,
                                   metadata={"index": ix.name, "k": k, "n_queries": nq, "batch_size": batch_size, "mode": mode})
                records.append(res.to_json())
                if best is None or (res.recall, res.queries_per_second) > (best.recall, best.queries_per_second):
                    best = res
        if best is None:
            return SearchResult(neighbors=np.empty((0, k), np.int64), distances=np.empty((0, k), np.float32), search_time_ms=0.0,
                                queries_per_second=0.0, recall=0.0, algorithm="", search_params=[], metadata={"dry_run": dry_run})
        best.metadata = dict(best.metadata, all_results=records)
        best.search_params = [r["search_params"][0] for r in records]
        return best


class CuvsB200Backend(HarnessMixin, BenchmarkBackend):
    """cuVS algorithms of the reference's benchmark configs, served by this library (GPU)."""

    requires_gpu = True
    ALGOS = ("cuvs_brute_force", "cuvs_ivf_flat", "cuvs_ivf_pq", "cuvs_cagra")

    def __init__(self, config: Dict[str, Any]):
        super().__init__(config)
        self._handles: Dict[str, Any] = {}
        self._res = None

    # ---- plugin lifecycle (base.py:452-480)
    def initialize(self) -> None:
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("cuvs_b200 benchmark backend needs a CUDA device (there is no CPU fallback)")
        from cuvs_b200.common import Resources
        self._res = Resources()

    def cleanup(self) -> None:
        self._handles.clear()
        self._res = None

    def _sync(self):
        if self._res is not None:
            self._res.sync()

    def _index_bytes(self, handle) -> int:
        return int(getattr(handle[0], "device_bytes", 0) or 0)

    # ---- algorithm dispatch: parameter names of cuvs_ann_bench_param_parser.h
    def _build_one(self, algo, metric, vectors, p):
        import torch
        if self._res is None:
            self.initialize()
        if algo not in self.ALGOS:
            raise ValueError(f"unknown algo {algo!r}; this backend serves {self.ALGOS}")
        ds = torch.from_numpy(np.ascontiguousarray(vectors)).cuda()
        n = ds.shape[0]
        if algo == "cuvs_brute_force":
            from cuvs_b200.neighbors import brute_force
            return (brute_force.build(ds, metric=metric, resources=self._res), ds)
        if algo == "cuvs_ivf_flat":
            from cuvs_b200.neighbors import ivf_flat
            ratio = float(p.get("ratio", 2))
            params = ivf_flat.IndexParams(n_lists=int(p.get("nlist", 1024)), metric=metric, kmeans_n_iters=int(p.get("niter", 20)),
                                          kmeans_trainset_fraction=min(1.0, 1.0 / ratio))
            return (ivf_flat.build(params, ds, resources=self._res), ds)
        if algo == "cuvs_ivf_pq":
            from cuvs_b200.neighbors import ivf_pq
            ratio = float(p.get("ratio", 2))
            params = ivf_pq.IndexParams(n_lists=int(p.get("nlist", 1024)), metric=metric, kmeans_n_iters=int(p.get("niter", 20)),
                                        kmeans_trainset_fraction=min(1.0, 1.0 / ratio), pq_dim=int(p.get("pq_dim", 0)),
                                        pq_bits=int(p.get("pq_bits", 8)))
            return (ivf_pq.build(params, ds, resources=self._res), ds)
        from cuvs_b200.neighbors import cagra
        gd = int(p.get("graph_degree", 64))
        params = cagra.IndexParams(metric=metric, graph_degree=gd, intermediate_graph_degree=int(p.get("intermediate_graph_degree", 2 * gd)))
        del n
        return (cagra.build(params, ds, resources=self._res), ds)

    def _search_batch(self, handle, algo, queries, k, sp, dataset):
        import torch
        index, ds = handle
        q = torch.from_numpy(np.ascontiguousarray(queries)).cuda()
        if algo == "cuvs_brute_force":
            from cuvs_b200.neighbors import brute_force
            d, i = brute_force.search(index, q, k, resources=self._res)
        elif algo == "cuvs_ivf_flat":
            from cuvs_b200.neighbors import ivf_flat
            d, i = ivf_flat.search(ivf_flat.SearchParams(n_probes=int(sp.get("nprobe", 20))), index, q, k, resources=self._res)
        elif algo == "cuvs_ivf_pq":
            from cuvs_b200.neighbors import ivf_pq, refine
            params = ivf_pq.SearchParams(n_probes=int(sp.get("nprobe", 20)),
                                         lut_dtype=_DTYPES[str(sp.get("smemLutDtype", "float"))],
                                         internal_distance_dtype=_DTYPES[str(sp.get("internalDistanceDtype", "float"))])
            rr = int(sp.get("refine_ratio", 1))
            if rr > 1:  # cuvs_ivf_pq_wrapper.h: search k * refine_ratio candidates, exact refine on the dataset
                cd, ci = ivf_pq.search(params, index, q, k * rr, resources=self._res)
                d, i = refine(ds, q, ci, k=k, metric=METRICS[getattr(dataset, "distance_metric", "euclidean")], resources=self._res)
            else:
                d, i = ivf_pq.search(params, index, q, k, resources=self._res)
        else:
            from cuvs_b200.neighbors import cagra
            params = cagra.SearchParams(itopk_size=int(sp.get("itopk", 64)), search_width=int(sp.get("search_width", 1)),
                                        max_iterations=int(sp.get("max_iterations", 0)))
            d, i = cagra.search(params, index, q, k, resources=self._res)
        self._res.sync()
        return d.cpu().numpy(), i.cpu().numpy().astype(np.int64)


def make_config_loader(algo_prefix: str = "cuvs_", backend_type: str = "cuvs_b200"):
    """A ConfigLoader for the reference's orchestrator (orchestrator/config_loaders.py:128-260): the shared base class loads the
    dataset YAML and expands the parameter grids; the two hooks below pick the algorithm YAML files this backend serves (names
    starting with `algo_prefix`, from the bundled config/algos directory plus `algorithm_configuration`) and turn every build
    combination into one IndexConfig / BenchmarkConfig — the same shape the OpenSearch loader produces
    (backends/opensearch.py:31-205).  Only available when the reference's cuvs_bench package is importable."""
    from cuvs_bench.orchestrator.config_loaders import BenchmarkConfig, ConfigLoader  # type: ignore
    from cuvs_bench.orchestrator.config_loaders import IndexConfig as RefIndexConfig  # type: ignore
    import cuvs_bench.backends as _ref_backends  # type: ignore

    class Loader(ConfigLoader):
        def __init__(self, config_path=None):
            self.config_path = os.fspath(config_path) if config_path is not None else os.path.join(os.path.dirname(os.path.realpath(_ref_backends.__file__)), "..", "config")  # the bundled config dir, as backends/opensearch.py:52-58 locates it

        @property
        def backend_type(self) -> str:
            return backend_type

        def _discover_algo_groups(self, dataset_conf, dataset, dataset_path, **kwargs):
            files = [f for f in self.gather_algorithm_configs(self.config_path, kwargs.get("algorithm_configuration"))
                     if os.path.basename(f).startswith(algo_prefix)]
            allowed_algos = [a.strip() for a in kwargs["algorithms"].split(",")] if kwargs.get("algorithms") else None
            allowed_groups = [g.strip() for g in kwargs["groups"].split(",")] if kwargs.get("groups") else None
            out = []
            for f in files:
                conf = self.load_yaml_file(f)
                name = conf.get("name", "")
                if allowed_algos and name not in allowed_algos:
                    continue
                for gname, gconf in conf.get("groups", {}).items():
                    if allowed_groups and gname not in allowed_groups:
                        continue
                    out.append((name, gname, gconf, {}))
            return out

        def _build_benchmark_configs(self, dataset_config, dataset_conf, dataset, dataset_path, expanded_groups, **kwargs):
            configs = []
            for algo_name, group_name, _conf, build_combos, search_combos, _meta in expanded_groups:
                for bp in build_combos or [{}]:
                    prefix = algo_name if group_name == "base" else f"{algo_name}_{group_name}"
                    label = ".".join([prefix] + [f"{k}{v}" for k, v in bp.items()])
                    ix = RefIndexConfig(name=label, algo=algo_name, build_param=bp, search_params=search_combos or [{}],
                                        file=os.path.join(dataset_path, dataset, "index", label))
                    configs.append(BenchmarkConfig(indexes=[ix], backend_config={"name": label, "algo": algo_name,
                                                                                 "requires_gpu": algo_prefix == "cuvs_"}))
            return configs

    return Loader


def register(name: str = "cuvs_b200") -> bool:
    """Register the backend AND its config loader with the reference's registries (backends/registry.py: register_backend,
    register_config_loader), after which `BenchmarkOrchestrator(backend_type=name).run_benchmark(...)` drives this library.
    False when cuvs_bench is absent."""
    if not HAVE_CUVS_BENCH:
        return False
    from cuvs_bench.backends.registry import get_registry  # type: ignore
    reg = get_registry()
    try:
        reg.register(name, CuvsB200Backend)
    except ValueError:
        pass  # already registered
    try:
        from cuvs_bench.backends.registry import register_config_loader  # type: ignore
        register_config_loader(name, make_config_loader("cuvs_", name))
    except Exception:  # noqa: BLE001 - older registry without config loaders, or already registered
        pass
    return True


def run_config(config: Dict[str, Any], dataset, k: int = 10, batch_size: int = 10000, mode: str = "throughput", backend=None):
    """One benchmark group in the shape of the reference's YAML configs (config/algos/*.yaml, `groups:` entries):
    {"name": "cuvs_ivf_pq", "groups": {"base": {"build": {"nlist": [1024], "pq_dim": [64]}, "search": {"nprobe": [20, 50]}}}}
    -> the Cartesian product of build parameters is built, every search combination is run; returns the JSON records
    (Google-Benchmark-compatible, base.py:240-257 / 310-337)."""
    import itertools
    algo = config["name"]
    backend = backend or CuvsB200Backend({"name": algo})
    backend.initialize()
    out = []
    try:
        for gname, group in config.get("groups", {"base": config}).items():
            bkeys = sorted(group.get("build", {}))
            skeys = sorted(group.get("search", {}))
            as_list = lambda v: v if isinstance(v, (list, tuple)) else [v]  # noqa: E731
            searches = [dict(zip(skeys, vals)) for vals in itertools.product(*[as_list(group["search"][x]) for x in skeys])] or [{}]
            indexes = []
            for vals in itertools.product(*[as_list(group["build"][x]) for x in bkeys]):
                bp = dict(zip(bkeys, vals))
                label = algo + "".join(f".{a}{b}" for a, b in bp.items())
                indexes.append(IndexConfig(name=f"{label}.{gname}", algo=algo, build_param=bp, search_params=searches, file=""))
            for ix in indexes:
                b = backend.build(dataset, [ix])
                out.append(b.to_json())
                s = backend.search(dataset, [ix], k=k, batch_size=batch_size, mode=mode)
                out.extend(s.metadata.get("all_results", []))
    finally:
        backend.cleanup()
    return out
