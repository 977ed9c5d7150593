"""The driver keeps a ~1.5 KB tail of bench.py's stdout and clips strings at 120 chars: the LAST line must be a
short, self-contained JSON object (round 1's 21 KB line was unreadable). Built here from a canned detail record."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def _canned():
    class A:
        vnum, nnz, feat_size, n_hidden, n_classes, fanout, batch_size = 10_000_000, 100_000_000, 600, 32, 60, "25,10", 6000
        config, tag, model, n_layers, partition, parts, scale = 2, "R-MAT", "gcn", 1, "hash", 0, 1.0
    kern = {"x" * 70: {"launches": 40, "avg_ms": 0.1773910, "achieved_gbs": 4213.123456789, "frac": 0.6430123456}}
    return {
        "value": 2725.123456789, "n_gpus": 8, "steps": 200, "warmup": 10, "ms_per_step": 0.36696123456,
        "config": {"workload": bench.workload_name(A), "cache_mode": "hbm20", "path": "engine", "kernel_timing": "y" * 900},
        "e2e": {"value": 2704.987654321, "h2d_bytes_per_step": 96000, "d2h_bytes_per_step": 188, "ms_per_step": 0.37},
        "gpu_launches": 5000, "gather_gbs": 2720.123456, "hit_rate": 1.0,
        "clocks": {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": ["sw_power_cap"], "power_w_max": 800.1, "samples": 20},
        "roofline": {"bound": "hbm", "kernel": "cache_aggregate_block0(agg_rows_tma_kernel,D=600)",
                     "achieved": 4213.123456789, "peak": 6551.7, "unit": "GB/s", "frac": 0.64301234, "traffic": 856512345,
                     "alg_bytes_per_launch": 747312345.6, "avg_ms": 0.17739123, "peak_source": "measured"},
        "cpu_baseline": {"value": 12.7428341, "unit": "minibatches/s", "cores": 16, "kind": "port", "sample": "32 full minibatches, oracle port, 16 threads",
                         "stage_s": {"sample": 1.0}},
        "parity_gate": "ok", "replicas_identical": True,
        "vtx20": {"value": 258.123456, "e2e": 257.123456, "gather_gbs": 200.123456, "hit_rate": 0.76123456,
                  "miss_frac_pcie": 0.87123456},
        "kernels": kern, "variants": {"hbm20": {"value": {"kernels": kern}}},
    }


def test_compact_line_is_short_and_complete():
    line = bench.compact_line(_canned())
    assert "\n" not in line and len(line) < 1150, len(line)
    d = json.loads(line)
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["unit"] == "minibatches/s"
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks", "parity_gate", "replicas_identical"):
        assert k in d, k
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["cpu_baseline"]) >= {"value", "unit", "cores", "kind", "sample"}
    assert d["e2e"]["h2d_bytes_per_step"] == 96000 and d["vtx20"]["hit_rate"] > 0

    def strings(o):
        if isinstance(o, str):
            yield o
        elif isinstance(o, dict):
            for x in o.values():
                yield from strings(x)
        elif isinstance(o, list):
            for x in o:
                yield from strings(x)
    assert max(len(s) for s in strings(d)) <= 120
    assert "kernels" not in d and "variants" not in d          # the detail lives in gpurun_out/bench_detail_n{N}.json


def test_compact_line_worst_case_strings_still_fit():
    c = _canned()
    c["cpu_baseline"]["sample"] = "s" * 300
    c["roofline"]["kernel"] = "k" * 200
    c["config"]["workload"] = "w" * 300
    line = bench.compact_line(c)
    assert len(line) < 1150 and json.loads(line)["value"] > 0


def test_every_config_resolves_and_names_itself():
    import sys
    for cfg in (1, 2, 3, 4, 5):
        old = sys.argv
        sys.argv = ["bench.py", "--config", str(cfg)]
        try:
            a = bench.parse_args()
        finally:
            sys.argv = old
        assert a.model in ("gcn", "gcn-pre", "sage") and len(a.fanout.split(",")) == (a.n_layers if a.model == "gcn-pre" else a.n_layers + 1)
        name = bench.workload_name(a)
        assert name.startswith("cfg%d " % cfg) and len(name) <= 120, name
        assert a.modes.split(",")[0].startswith("hbm") and a.modes.split(",")[1].startswith("vtx")


def test_compact_line_without_optional_parts():
    c = _canned()
    c["cpu_baseline"] = None
    c.pop("vtx20")
    c["clocks"] = None
    d = json.loads(bench.compact_line(c))
    assert d["cpu_baseline"] is None and "vtx20" not in d and d["clocks"]["reasons"] == []


def test_dominant_kernel_skips_the_sampler_chain_and_the_all_reduce():
    k = {"sample(all kernels of one pg_sample)": {"bound": "hbm", "avg_ms": 0.26},
         "allreduce+adam(allreduce_adam_kernel)": {"bound": "hbm", "avg_ms": 0.20},
         "gather_miss(rows_bulk_kernel)": {"bound": "pcie", "avg_ms": 3.0},
         "cache_aggregate_blocks0..1(agg_rows_tma_kernel,D=600)": {"bound": "hbm", "avg_ms": 0.0975},
         "gather_hit(rows_ldg_kernel)": {"bound": "hbm", "avg_ms": 0.0247}}
    assert bench.dominant_kernel(k).startswith("cache_aggregate")
