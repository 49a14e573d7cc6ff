"""CPU: third-party arithmetic that is NOT under /root/reference (OpenCV, un-pinned in requirements.txt; the image has
4.13.0) restated in the oracle, pinned by running the installed library side by side."""
import numpy as np
import pytest

from conftest import assert_same

cv2 = pytest.importorskip("cv2")


def test_rgb2gray(orc):
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    assert_same(orc.rgb2gray(img), cv2.cvtColor(img, cv2.COLOR_RGB2GRAY), "RGB2GRAY")
    assert_same(orc.bgr2gray(img), cv2.cvtColor(img, cv2.COLOR_BGR2GRAY), "BGR2GRAY")


@pytest.mark.parametrize("shape,pads", [((37, 50, 3), (4, 5, 3, 3)), ((16, 16), (0, 0, 0, 0)), ((5, 9), (7, 8, 7, 9)),
                                        ((375, 1242, 3), (4, 5, 3, 3))])
def test_pad_reflect(orc, shape, pads):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    t, b, l, r = pads
    assert_same(orc.pad_reflect(img, t, b, l, r), cv2.copyMakeBorder(img, t, b, l, r, cv2.BORDER_REFLECT), "BORDER_REFLECT")


@pytest.mark.parametrize("seed", range(6))
def test_filter_speckles(orc, seed):
    rng = np.random.default_rng(seed)
    H, W = 60 + seed, 90
    base = np.kron(rng.integers(0, 90, (H // 6 + 1, W // 6 + 1)), np.ones((6, 6)))[:H, :W]
    img = (base + rng.integers(0, 12, (H, W))).astype(np.uint8)
    img[rng.random((H, W)) < 0.15] = 0
    want = img.copy()
    cv2.filterSpeckles(want, 0, 200, 10)
    assert_same(orc.filter_speckles_u8(img, 0, 200, 10), want, "filterSpeckles(0, 200, 10)")
    want = img.copy()
    cv2.filterSpeckles(want, 0, 20, 3)
    assert_same(orc.filter_speckles_u8(img, 0, 20, 3), want, "filterSpeckles(0, 20, 3)")
