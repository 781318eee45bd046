"""CPU: the host side of the ResNet harness (optimal_conv_b200/resnet.py) -- the layer sequence of testResNet_crop_sparse,
the weight-file layout, the strided layers' kernel split and the compare_final.py acceptance check on files."""
import os

import numpy as np

from optimal_conv_b200 import hostprep as hp, resnet


def test_layer_sequence_is_the_reference_s():
    """test.go:112-262: 7 + 1 + 5 + 1 + 5 bootstrapped layers and the final FC; packing densities 2,1,3,2,4; weights w0..w18"""
    L = resnet.layer_specs(20, 3)
    assert len(L) == 20 and [s["kind"] for s in L].count("Conv_sparse") == 17 and [s["kind"] for s in L].count("StrConv_sparse") == 2
    assert [s["log_sparse"] for s in L[:-1]] == [2] * 7 + [1] + [3] * 5 + [2] + [4] * 5
    assert [s["w"] for s in L[:-1]] == list(range(19)) and L[0]["real_ib"] == 3 and L[1]["real_ib"] == 16
    assert [(s["in_wid"], s["kp_wid"]) for s in (L[0], L[7], L[8], L[13], L[14])] == [(32, 31), (32, 15), (16, 15), (16, 7), (8, 7)]
    assert L[-1] == dict(name="final_fc", kind="final", in_wid=8, kp_wid=7, ker_wid=7, real_ib=64, real_ob=10, norm=16)
    assert len(resnet.layer_specs(8, 3)) == 3 + 1 + 1 + 1 + 1 + 1 and resnet.layer_specs(20, 3, cf100=True)[-1]["real_ob"] == 100
    assert [s["pow"] for s in resnet.layer_specs(20, 3, cf100=True)[:-1]] == [5.0] * 18 + [7.0]


def test_weight_files_and_kernel_split(tmp_path):
    d = str(tmp_path) + "/"
    L = resnet.layer_specs(20, 3)
    for s in (L[0], L[7]):
        ker, a, b = resnet.synthetic_weights(s, 3, 5 + s["w"])
        for f, v in zip(hp.weight_files(d, s["w"]), (ker, a, b)):
            hp.write_txt(f, v)
        got = resnet.load_weights(d, s, 3)
        assert all(np.array_equal(x, y) for x, y in zip(got, (ker, a, b)))   # shortest-digits text round-trips doubles
    # "StrConv_sparse" (eval.go:347-366): ker_in_0[k][i][j] = ker_in[k][i][2j], ker_in_1 = [2j+1]; bn vectors likewise
    s = L[7]
    ker, a, b = resnet.synthetic_weights(s, 3, 1)
    (k0, a0, b0), (k1, a1, b1) = resnet.split_by_output_parity(ker, a, b, 9, s["real_ib"], s["real_ob"])
    ib, ob = s["real_ib"], s["real_ob"]
    for k in (0, 4, 8):
        for i in (0, 7, 15):
            for j in (0, 5, 15):
                assert k0[k * ib * ob // 2 + i * ob // 2 + j] == ker[k * ib * ob + i * ob + 2 * j]
                assert k1[k * ib * ob // 2 + i * ob // 2 + j] == ker[k * ib * ob + i * ob + 2 * j + 1]
    assert np.array_equal(a0, a[0::2]) and np.array_equal(b1, b[1::2])
    # final layer: the FC kernel repeated at every tap of the 7x7 window, bn_a = 1 / raw_wid^2 (test.go:296-313)
    fc = np.arange(640, dtype=float)
    hp.write_txt(d + "final-fckernel.csv", fc)
    hp.write_txt(d + "final-fcbias.csv", np.arange(10, dtype=float))
    ker, a, b = resnet.load_weights(d, L[-1], 3)
    assert ker.shape == (49 * 640,) and np.array_equal(ker[640 * 3:640 * 4], fc) and np.allclose(a, 1 / 49) and b[3] == 3


def test_compare_logits_is_compare_final(tmp_path):
    """compare_final.py:8-64 on a synthetic directory pair: accordance, precisions, missing files skipped, tolerance"""
    plain_dir, enc_dir = str(tmp_path / "plain"), str(tmp_path / "enc")
    os.makedirs(plain_dir), os.makedirs(enc_dir)
    rng = np.random.default_rng(3)
    n, c = 100, 10
    plain = rng.normal(size=(n, c))
    labels = np.argmax(plain, axis=1).astype(float)
    labels[5] = (labels[5] + 1) % c                                   # the plain model is wrong on image 5
    hp.write_txt(os.path.join(plain_dir, "plain_prediction_%d.csv" % n), plain.reshape(-1))
    hp.write_txt(os.path.join(plain_dir, "test_labels_%d.csv" % n), labels)
    for i in range(0, n, 3):                                          # the encrypted run covered every third image
        e = plain[i] + rng.normal(scale=1e-3, size=c)
        if i == 9:
            e[np.argmin(e)] = 10.0                                    # ... and got one of them wrong
        hp.write_txt(os.path.join(enc_dir, "class_result_ker3_%d.csv" % i), e)
    r = resnet.compare_logits(enc_dir, plain_dir, 3, n)
    assert r["compared"] == 34 and r["accordance"] == 33 and r["enc_precision"] == 33 and r["plain_precision"] == 34
    assert r["max_abs_diff"] > 5
    import pytest
    with pytest.raises(AssertionError):
        resnet.compare_logits(enc_dir, plain_dir, 3, n, tol=0.05)
