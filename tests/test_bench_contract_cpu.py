"""The reference arm of bench.py runs on the host only (the oracle port on CPU): its JSON line must carry the keys the
driver's contract names.  (The GPU arm's line is produced on the B200 box; profiles/r02_bench_final.json is its latest record.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1",
                        "--cpu-sample", "8"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "smpl_bodies_per_sec" and line["unit"] == "bodies/s"
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["sample"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert "workload" in line["config"] and "model" not in line["config"]


def test_committed_gpu_bench_line_has_the_contract_keys():
    """the last GPU bench line recorded under profiles/ (written on the B200 box by the default `python bench.py`)"""
    line = json.loads(open(os.path.join(ROOT, "profiles", "r02_bench_final.json")).read().strip().splitlines()[-1])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert k in line, k
    assert line["vs_baseline"] is None and line["gpu_launches"] > 0 and line["warmup"] >= 3
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert set(("value", "unit", "cores", "kind", "sample")) <= set(line["cpu_baseline"])
    assert set(("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step")) <= set(line["e2e"])
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # round-2 legs: the loop with the extractors' MLP fused into sampling, the training step, the other BASELINE configs
    rd = line["with_reduce_dim"]
    assert rd["parity_ref_features_rel"] <= 1e-4
    assert rd["fused_nchw"]["ms_per_step"] < rd["two_step_nchw"]["ms_per_step"]
    assert rd["fused_channels_last"]["ms_per_step"] < rd["two_step_channels_last"]["ms_per_step"]
    assert line["train_step"]["ms_forward_loss_backward"] < line["train_step"]["ms_eager_autograd_same_gpu"]
    assert max(line["train_step"]["grad_rel_diff_vs_eager"].values()) <= 1e-4
    assert set(("smpl_sweep_65536", "eval_pass_35515")) <= set(line["other_configs"])
    assert line["parity"]["verts_m"] <= 1e-5 and line["parity"]["kp2d_px"] <= 1e-3 and line["parity"]["sampled_rel"] <= 1e-4
    # last session's legs: end to end with host-resident feature levels gathered in place (headline) beside the all-copies
    # pass, both tensor-core arithmetics of the fused SMPL kernel, the share of the whole loop, every BASELINE config
    e2e, copy_all = line["e2e"], line["e2e_copy_all"]
    assert e2e["results_identical_to_device_resident_step"] is True and e2e["host_resident_levels"]
    assert e2e["value"] > copy_all["value"] and e2e["h2d_bytes_per_step"] < copy_all["h2d_bytes_per_step"]
    assert e2e["host_input_bytes_per_step"] == copy_all["h2d_bytes_per_step"]
    assert line["e2e_host_gather_channels_last"]["sampled_rel_vs_nchw_step"] <= 1e-4
    modes = line["pose_blend_modes"]
    for name in ("bf16x3", "3xtf32"):
        assert modes[name]["fused_kernel"] is True and modes[name]["verts_max_err_vs_fp64_m"] <= 1e-5
        assert 0.0 < modes[name]["tensor_frac"] < 1.0
    assert modes["3xtf32"]["verts_max_err_vs_fp64_m"] < modes["bf16x3"]["verts_max_err_vs_fp64_m"]
    wl = line["whole_loop"]
    assert wl["ms_whole_loop_new"] < wl["ms_whole_loop_old"] and wl["hot_path_share_new"] < wl["hot_path_share_old"]
    assert wl["verts_new_vs_old_m"] <= 1e-5
    assert set(("smpl_b64", "maf_sampling_1024x431")) <= set(line["other_configs"])
