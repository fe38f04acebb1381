"""bench.py, host side (no GPU): the --impl reference arm prints a line the driver can pair with the GPU arm's --
same metric, unit, direction and, key for key, the same `config` -- and `static_config` reproduces the `config` the GPU arm
printed on the B200 boxes (committed driver-format lines under profiles/).  The CPU arm times oracle/ (allowed: it is the
reference arm of the bench, never the product path)."""
import json
import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402

MOVED = ("transport", "fused", "strong_scaling", "slab_bit_identical")     # run-time keys, now under `roofline`


def _args(**kw):
    a = dict(workload="c4_sc_d3q19_512", size="", scaling="weak", fused=1, gpus=1, steps=1, warmup=3)
    a.update(kw)
    return types.SimpleNamespace(**a)


def _committed_line(name):
    for ln in open(os.path.join(ROOT, "profiles", name)):
        if ln.startswith('{"metric'):
            return json.loads(ln)
    raise AssertionError("no bench line in " + name)


@pytest.mark.parametrize("name,world", [("bench_r2_final3_n1.json", 1), ("bench_r2_final2_n2.json", 2), ("bench_r2_final2_n8.json", 8)])
def test_static_config_is_what_the_gpu_arm_printed(name, world):
    pkg = bench.entry.load_package()
    printed = {k: v for k, v in _committed_line(name)["config"].items() if k not in MOVED}
    assert bench.static_config(pkg, _args(gpus=world), world) == printed


def test_static_config_strong_and_size():
    pkg = bench.entry.load_package()
    c = bench.static_config(pkg, _args(scaling="strong", gpus=8), 8)
    assert c["lattice_global"] == [512, 512, 512] and c["lattice_per_gpu"] == [64, 512, 512]
    c = bench.static_config(pkg, _args(workload="c3_hcz_d2q9_full", scaling="strong", gpus=8), 8)
    assert c["lattice_global"] == [2048, 8194, 1] and c["lattice_per_gpu"] == [256, 8194, 1]
    c = bench.static_config(pkg, _args(size="128x64"), 1)
    assert c["lattice_per_gpu"] == [128, 64, 1] and c["lattice_global"] == [128, 64, 1]
    c = bench.static_config(pkg, _args(workload="c1_sc_d2q9_256"), 1)
    assert "FITS in the 126 MB L2" in c["l2_policy"]


@pytest.mark.parametrize("world,rank", [(1, 0), (2, 0), (2, 1)])
def test_reference_arm_line(monkeypatch, capsys, world, rank):
    real = bench.cpu_baseline
    monkeypatch.setattr(bench, "cpu_baseline", lambda P, key, threads=0, target_s=0.0: real(P, key, threads, target_s=0.2))
    monkeypatch.setattr(bench, "reference_functor_baseline", lambda key, same_lattice=None: None)   # timed in its own tests
    monkeypatch.setenv("WORLD_SIZE", str(world))
    a = _args(gpus=world)
    bench.run_reference_arm(a, rank)
    out = capsys.readouterr().out.strip()
    if rank != 0:
        assert out == ""            # the other ranks exit without work
        return
    line = json.loads(out)
    gpu = _committed_line("bench_r2_final3_n1.json")
    for k in ("metric", "unit", "higher_is_better", "dtype", "data", "scaling"):
        assert line[k] == gpu[k], k
    assert line["impl"] == "reference" and line["n_gpus"] == world and line["steps"] == 1
    assert line["config"] == bench.static_config(bench.entry.load_package(), a, world)
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == bench.host_threads() and cb["value"] == line["value"]
    assert cb["sample_lattice"] == [96, 96, 96] and "96x96x96" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "MLUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "c4_hcz_d3q19_512" in cb["also"]


def test_host_threads_ignores_omp_num_threads(monkeypatch):
    monkeypatch.setenv("OMP_NUM_THREADS", "1")      # what torchrun exports to every rank
    assert bench.host_threads() == len(os.sched_getaffinity(0))


class _FakeLat:
    def close(self):
        pass


class _FakeCtx:
    """stands in for bench.Ctx so that main()'s assembly of the JSON line runs without a device"""
    world, rank = 1, 0

    def __init__(self, a):
        self.rank, self.local_rank = type(self).rank, 0
        self.world = type(self).world
        self.pkg = bench.entry.load_package()
        self.dev = types.SimpleNamespace(index=0)
        self.dist = types.SimpleNamespace(destroy_process_group=lambda: None)

    def barrier(self):
        pass


def _fake_measure(cx, workload, scaling, steps, warmup, fused=1, size="", solo=False, sample_clocks=False, keep=False, ktiming=True):
    key, sz, desc = bench.WORKLOADS[workload]
    world = 1 if solo else cx.world
    nxl, ny, nz = sz
    nxg = nxl * world if (scaling == "weak" or world == 1) else nxl
    if nxg == nxl and world > 1:
        nxl = nxl // world
    res = {"workload": workload, "description": desc, "scaling": scaling, "n_gpus": world, "lattice_per_gpu": [nxl, ny, nz],
           "lattice_global": [nxg, ny, nz], "steps": steps, "ms_per_step": 1.0, "mlups": 1000.0 * world, "mass": 1.0, "gpu_launches": steps,
           "transport": "peer" if world > 1 else None, "kernel": "k", "kernel_ms": 1.0, "kernel_launches_sampled": steps,
           "algorithmic_bytes_per_lu": 305, "lattice_updates_per_launch": nxl * ny * nz, "achieved_gbs": 5000.0, "peak_gbs": 6545.0,
           "peak_source": "test", "frac": 0.76, "step_frac_of_peak": 0.75, "bytes_per_gpu": nxl * ny * nz * 39 * 8, "clocks": None}
    return (res, _FakeLat(), None, None) if keep else res


@pytest.mark.parametrize("world", [1, 2])
def test_gpu_arm_line_assembly(monkeypatch, capsys, world):
    """main() with the device work stubbed out: the line parses, `config` is static_config (so equal to the reference arm's),
    the run-time keys sit under `roofline`."""
    _FakeCtx.world = world
    monkeypatch.setattr(bench, "Ctx", _FakeCtx)
    monkeypatch.setattr(bench, "measure", _fake_measure)
    monkeypatch.setattr(bench, "run_e2e", lambda *a, **k: {"value": 1.0, "unit": "MLUPS", "h2d_bytes_per_step": 1, "d2h_bytes_per_step": 1})
    monkeypatch.setattr(bench, "slab_bit_identical", lambda cx: {"sc_d3q19": {"bit_identical": True}})
    monkeypatch.setattr(bench, "cpu_baseline", lambda *a, **k: {"value": 1.0, "kind": "port"})
    monkeypatch.setattr(bench, "reference_functor_baseline", lambda *a, **k: None)
    monkeypatch.setattr(bench, "pulsatile_extra", lambda *a, **k: {"mlups": 1.0})
    monkeypatch.setenv("WORLD_SIZE", str(world))
    monkeypatch.setenv("RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", str(world), "--steps", "20", "--warmup", "5"])
    bench.main()
    line = json.loads(capsys.readouterr().out.strip())
    assert line["config"] == bench.static_config(bench.entry.load_package(), _args(gpus=world), world)
    assert line["n_gpus"] == world and line["value"] == 1000.0 * world
    rf = line["roofline"]
    assert rf["fused"] == 1 and rf["transport"] == ("peer" if world > 1 else None)
    if world == 1:
        assert set(rf["also"]) == {"c4_hcz_d3q19_512", "c3_hcz_d2q9_full", "sc_d2q9_8192", "c1_sc_d2q9_256", "c2_hcz_d2q9_256", "c5_pulsatile_1024"}
    else:
        assert set(rf["strong_scaling"]) == {"c4_sc_d3q19_512", "c4_hcz_d3q19_512", "c3_hcz_d2q9_full"}
        assert all(abs(v["strong_efficiency"] - 1.0) < 1e-12 for v in rf["strong_scaling"].values())
        assert rf["slab_bit_identical"]["sc_d3q19"]["bit_identical"] is True
    assert not any(k in line["config"] for k in MOVED)


def test_pulsatile_and_yl2d_configs_are_what_the_gpu_arm_printed():
    a = _args(workload="c5_pulsatile_1024")
    assert bench.pulsatile_config(a, 1, 1024) == _committed_line("bench_r2_puls_1024.json")["config"]
    c = bench.yl2d_config(_args(workload="yl2d_8192"), 1, 8192)
    assert c["lattice_per_gpu"] == [8192, 8192, 1] and "19.3 GB" in c["l2_policy"]


def test_traffic_sources_cite_committed_summaries():
    table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    for key, e in table.items():
        traffic, src = bench.ncu_traffic(key)
        assert traffic == e["dram_bytes_read"] + e["dram_bytes_write"]
        assert os.path.exists(os.path.join(ROOT, src.split(" ")[0])), src       # the file the line points the reader to exists
    assert bench.ncu_traffic("no_such_workload") == (None, None)


def test_pulsatile_cpu_arm_is_the_untouched_reference_from_the_open_vessel():
    from _oracle import ref_binary
    if not ref_binary("ref_pulsatile"):
        pytest.skip("oracle/_ref/ref_pulsatile not built (no /root/reference here)")
    cb = bench.pulsatile_cpu_baseline(64, target_s=0.2)
    assert cb["kind"] == "reference" and cb["cores"] == 1 and cb["value"] > 0 and "open vessel" in cb["sample"]
