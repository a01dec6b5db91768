"""Host-side contract of bench.py (no GPU): the workload table matches BASELINE.json, both arms print the same `config`
object, and the algorithmic bytes per solve are the figures SURVEY section 8(d) states."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_algorithmic_bytes_are_the_survey_figures():
    w = bench.WORKLOADS
    assert bench.algorithmic_bytes(w["cfg2"]) == 6496
    assert [bench.algorithmic_bytes(w[f"cfg4-{n}"]) for n in (20, 40, 80)] == [4416, 8576, 16896]
    assert bench.algorithmic_bytes(w["cfg3"]) == 58080
    assert bench.algorithmic_bytes(w["cfg5"]) == 7936


def test_workloads_are_the_baseline_configs():
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        cfgs = json.load(f)["configs"]
    w = bench.WORKLOADS
    assert "N=30" in cfgs[1] and "4096" in cfgs[1] and w["cfg2"]["N"] == 30 and w["cfg2"]["batch"] == 4096
    assert "13 sigma points" in cfgs[2] and "1024" in cfgs[2] and w["cfg3"]["n_samples"] == 13 and w["cfg3"]["batch"] == 1024
    assert w["cfg3"]["closed_loop"] and "closed-loop" in cfgs[2]
    assert "8192" in cfgs[3] and all(w[f"cfg4-{n}"]["batch"] == 8192 and w[f"cfg4-{n}"]["track"] == "lvms" for n in (20, 40, 80))
    assert "16384" in cfgs[4] and w["cfg5"]["batch"] * 8 == 16384 and w["cfg5"]["track"] == "modena"
    for name, spec in w.items():
        ocp = bench.make_ocp(spec)
        assert ocp.N == spec["N"] and abs(ocp.dt - 0.08) < 1e-15 and ocp.variant == spec["variant"]
        # one dict for both arms: a pure function of (workload, world size)
        assert bench.workload_config(name, 4) == bench.workload_config(name, 4)
        assert bench.workload_config(name, 1)["workload"] == spec["label"]


def test_reference_arm_prints_the_same_config_object(capsys):
    """`--impl reference` on the smallest sample that still exercises the path: the JSON line carries `impl`, the e2e and
    cpu_baseline objects the contract names and the very `config` object the GPU arm prints."""
    class A:
        workload, gpus, steps, warmup = "cfg4-20", 1, 1, 0
    old = bench.cpu_sample_size
    bench.cpu_sample_size = lambda name: 16
    try:
        bench.run_reference(A)
    finally:
        bench.cpu_sample_size = old
    d = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["config"] == bench.workload_config("cfg4-20", 1)
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["unit"] == "solves/s" and d["higher_is_better"] is True and d["gpu_launches"] == 0
