"""The driver's bench contract, CPU side: `bench.py --impl reference` prints ONE JSON line with the agreed keys
(shrunk workload so that the test takes seconds; the GPU arm's line is exercised on the B200 by the driver)."""
import argparse
import json


def test_reference_arm_prints_the_contract_line(capsys, monkeypatch):
    import bench
    for k, v in dict(E_LAYERS=1, D_LAYERS=1, T_FRAMES=32, N_TEXT=4).items():
        monkeypatch.setattr(bench, k, v)
    monkeypatch.delenv("RANK", raising=False)
    bench.run_reference_arm(argparse.Namespace(gpus=1, steps=1, warmup=1))
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
                "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "clips/s" and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert cb["train_step_value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0


def test_non_zero_ranks_of_the_reference_arm_stay_silent(capsys, monkeypatch):
    import bench
    monkeypatch.setenv("RANK", "1")
    bench.run_reference_arm(argparse.Namespace(gpus=2, steps=1, warmup=1))
    assert capsys.readouterr().out.strip() == ""
