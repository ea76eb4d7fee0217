"""The roofline bookkeeping of ``bench.py`` uses the symmetric-aware algorithmic work of SURVEY 8(d); pin the
formulas against the figures the survey states for the BASELINE configs."""

import bench


def test_gram_flops_match_the_survey_figures():
    R = 1280  # c2: C N = 10 * 128
    conv = [4800, 64, 55296, 96, 110592, 128]  # weights and biases of the three conv layers, D_p = 170 976
    total = sum(bench.algorithmic_work("gram_dense_accum", [[R, R], [R, d]], 4)[1] for d in conv)
    assert sum(conv) == 170976 and total == R * (R + 1) * 170976
    assert abs(total - 2.80e11) / 2.80e11 < 0.01  # "2.80e11 (conv weights+biases dense)"
    kind, cross = bench.algorithmic_work("gram_cross_accum", [[R, 128], [R, 110592]], 4)
    assert kind == "tensor" and cross == 2 * R * 128 * 110592
    # c4: one hidden layer, R = 5120, N = 512, 4096 x 4096
    R, N = 5120, 512
    _, lin = bench.algorithmic_work("gram_linear_accum", [[R, R], [10, N, 4096], [N, 4096]], 4)
    assert lin == R * (R + 1) * 4096 + N * (N + 1) * 4096 + R * (R + 1) // 2
    assert abs(3 * lin - 3.3e11) / 3.3e11 < 0.03  # "c4 ~ 3.3e11 (dominated by R(R+1) 4096 x 3)"


def test_bandwidth_bound_kernels_count_each_operand_once():
    K, R, D, es = 10, 1280, 110592, 4
    kind, nbytes = bench.algorithmic_work("backtransform_dense", [[K, R], [R, D]], es)
    assert kind == "hbm" and nbytes == es * (R * D + K * R + K * D)
    kind, nbytes = bench.algorithmic_work("sqrt_backprop_elementwise", [[10, 128, 64, 28, 28], [128, 64, 28, 28]], es)
    assert kind == "hbm" and nbytes == es * 2 * 10 * 128 * 64 * 28 * 28
    assert bench.algorithmic_work("syevj", [[R, R]], es) == (None, 0)  # latency-bound: no roofline figure
    # the conv factor emit reads S and the layer input and WRITES the factor (the written
    # bytes used to be left out)
    S, X, Vt = [10, 128, 96, 12, 12], [128, 64, 14, 14], [10, 128, 96, 64, 3, 3]
    kind, nbytes = bench.algorithmic_work("v_emit_conv2d", [S, X], es, [Vt])
    numel = lambda shape: int(__import__("math").prod(shape))  # noqa: E731
    assert kind == "hbm" and nbytes == es * (numel(S) + numel(X) + numel(Vt))
    # structured Linear back-transform: K (C N out + N out in) multiply-adds
    kind, flops = bench.algorithmic_work("backtransform_linear", [[10, 5120], [10, 512, 4096], [512, 4096]], es)
    assert kind == "tensor" and flops == 2 * 10 * 512 * 4096 * (10 + 4096)


def test_workloads_are_the_baseline_configs():
    assert sorted(bench.WORKLOADS) == ["c1", "c2", "c3", "c4", "c5"]
    for key, n, classes in (("c1", 32, 10), ("c2", 128, 10), ("c3", 128, 100), ("c4", 512, 10), ("c5", 1024, 10)):
        w = bench.WORKLOADS[key]
        assert (w["n"], w["classes"]) == (n, classes)
    assert bench.WORKLOADS["c3"]["sub_ggn"] == list(range(32)) and bench.WORKLOADS["c3"]["mc"] == 1
    params = lambda m: sum(p.numel() for p in m.parameters())  # noqa: E731
    assert params(bench.mlp_c1()) == 52650            # D of c1
    assert params(bench.cnn_3c3d()) == 895210         # c2
    assert params(bench.allcnnc()) == 1387108         # c3
    assert params(bench.mlp_c4()) == 36818954         # c4
