"""1920x1080 output fitting (render.py:98-105): the host coefficient tables against PIL itself (CPU), and the device
kernel against PIL bit-for-bit (GPU)."""
import numpy as np
import pytest

from oracle import ops_oracle as OO


@pytest.mark.parametrize("in_hw,out_hw", [((64, 114), (68, 120)), ((33, 57), (35, 60)), ((40, 40), (23, 31)),
                                          ((1024, 1824), (1080, 1920))])
def test_pillow_coefficient_tables_reproduce_pil(in_hw, out_hw):
    import PIL.Image

    from maua_stylegan2_b200.render import pillow_bilinear_coeffs

    rng = np.random.Generator(np.random.PCG64(in_hw[0]))
    img = rng.integers(0, 256, in_hw + (3,), dtype=np.uint8)
    bx, kx = pillow_bilinear_coeffs(in_hw[1], out_hw[1])
    by, ky = pillow_bilinear_coeffs(in_hw[0], out_hw[0])
    mine = OO.resample_fixed_point(img, bx, kx, by, ky)
    ref = np.array(PIL.Image.fromarray(img).resize((out_hw[1], out_hw[0]), PIL.Image.BILINEAR))
    assert np.array_equal(mine, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(3, 1024, 2048, 3), (2, 2048, 1024, 3), (2, 512, 512, 3)])
def test_fit_frames_kernel_bit_exact_vs_pil(shape):
    import torch

    from maua_stylegan2_b200.render import fit_frames

    rng = np.random.Generator(np.random.PCG64(1))
    frames = rng.integers(0, 256, shape, dtype=np.uint8)
    frames[0, :, :, :] = (np.arange(shape[2])[None, :, None] * 7 + np.arange(shape[1])[:, None, None] * 3) % 256  # ramps
    out = fit_frames(torch.from_numpy(frames).cuda(), 1920 if shape[2] >= shape[1] else 1080).cpu().numpy()
    ref = OO.fit_output(frames)
    assert out.shape == ref.shape
    assert np.array_equal(out, ref)
