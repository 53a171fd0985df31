"""bench.py's numpy generators must be the same data the oracle's C generators produce."""
import numpy as np

import bench
import oracle


def test_noise_matches_oracle():
    x = bench.synth_noise(5, 3, 100, 777)
    for c in range(3):
        assert np.array_equal(x[c], oracle.gen_noise(5 + c, 100, 777))


def test_irs_match_oracle():
    h = bench.synth_irs(2, 3, 4, 4800, chunk=2)
    for c in range(3):
        ref = oracle.gen_ir(2 + c, 4, 4800)
        assert np.max(np.abs(h[c] - ref)) <= 1e-9  # f64 sum order may differ in the last ulp of the norm


def test_both_arms_quote_the_same_config_dict():
    """the reference arm's `config` must be the CUDA arm's (the driver compares them); what each arm timed per step
    lives in `cpu_baseline.sample` / `tuning`"""
    import argparse
    a = argparse.Namespace(channels=4096, block=512, ir_seconds=2.0)
    d = bench.config_dict(a)
    assert set(d) == {"workload", "channels_per_gpu", "block", "ir_taps", "l2"} and d["ir_taps"] == 96000
    import inspect
    src = inspect.getsource(bench.run_reference) + inspect.getsource(bench.run_b200)
    assert src.count('"config": config_dict(args)') == 2


def test_traffic_is_only_quoted_for_the_kernels_it_was_captured_from():
    import json
    h = bench.kernel_source_hash()
    assert len(h) == 16 and h == bench.kernel_source_hash()
    t = json.loads((bench.ROOT / "profiles" / "fused_traffic.json").read_text())
    assert "kernel_source_sha256_16" in t and t["channels"] == 4096
    # 6.35 GB per launch: within 1 % of the algorithmic bytes of the fused kernel at 4096 channels
    S, K, B = 188, 513, 512
    algorithmic = 4096 * (16 * (S - 1) * K + 16 * K + 16 * B)
    assert abs(t["dram_bytes_per_launch"] / algorithmic - 1.0) < 0.01


def test_numa_binding_is_a_noop_on_a_single_node_box():
    info = bench.bind_to_gpu_numa_node(0)  # no GPU here: must not raise, must not bind
    assert info["bound"] is False
