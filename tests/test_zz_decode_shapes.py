"""The general packet-decode kernel has two shapes: 128 threads per frame, and one warp per frame for
launches with more frames than the first shape has CTAs.  The parity cases are small, so on their own they only ever
reach the 128-thread shape; B2_VIT_MODE=2 forces the one-warp shape, here on every block code of the path
(Hamming(12,8), Golay(24,12), both at once, conv r1/2 K=7), clean and noisy.  (Sorted last on purpose: it re-runs
cases the earlier files already cover through the default shape.)"""
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c2_8ch_h128", "golay_outer_h128", "m256_qam256_golay", "noisy_30dB", "n5_m96_v27"])
def test_one_warp_shape_of_the_decode_kernel_on_every_code(name, monkeypatch):
    import test_gpu_parity as T
    monkeypatch.setenv("B2_VIT_MODE", "2")
    case = T.CASES[name]
    x = T.make_input(case)
    fo, po, _ = T.run_oracle(case, x)
    fg, pg, _ = T.run_gpu(case, x)
    assert len(fo) > 0 and int(fo["payload_valid"].sum()) > 0
    T.assert_frames_equal(fo, po, fg, pg)
