"""The GPU comparator (oracle/gpu_reference.py = the reference's own CUDA path: its compiled op/ extension + cuDNN grouped
convs) against the CPU oracle, and the product against it — north_star: "outputs must match the reference op/ CUDA path
... within 1e-3 relative fp32 for generator activations on identical latents/noise"."""
import numpy as np
import pytest
import torch

from oracle import stylegan2_oracle as O
from tests.util import make_generator, rel_err

pytestmark = pytest.mark.gpu


def test_product_matches_reference_cuda_path():
    from oracle.build_ref import load_ref

    if load_ref("upfirdn2d_ref") is None or load_ref("fused_ref") is None:
        pytest.skip("oracle/_ref is not built: the reference's CUDA ops are unavailable on this box")
    from oracle import gpu_reference as GR

    size, cm, seed, b = 256, 2, 0, 2
    g, sd = make_generator(size, cm, seed, "tc")
    _, num_layers, n_latent = O.layout(size)
    rng = np.random.Generator(np.random.PCG64(31))
    latent = torch.from_numpy(rng.standard_normal((b, n_latent, 512)).astype(np.float32)) * 0.5
    noise = [torch.from_numpy(rng.standard_normal((b, 1, 2 ** ((l + 5) // 2), 2 ** ((l + 5) // 2))).astype(np.float32))
             for l in range(num_layers)]
    psi = torch.tensor([0.7, 1.0])
    tl = torch.from_numpy(rng.standard_normal((1, 512)).astype(np.float32)) * 0.1
    sd_cuda = {k: v.cuda() for k, v in sd.items()}
    with torch.no_grad():
        cpu_img, cpu_acts = O.generator_forward(sd, size, latent, noise, psi, tl, channel_multiplier=cm)
        ref_img, ref_acts = GR.forward(sd_cuda, size, latent.cuda(), [n.cuda() for n in noise], psi.cuda(), tl.cuda(), cm,
                                       allow_tf32=False)
        tf_img, tf_acts = GR.forward(sd_cuda, size, latent.cuda(), [n.cuda() for n in noise], psi.cuda(), tl.cuda(), cm,
                                     allow_tf32=True)
        g.truncation_latent = tl.cuda()
        img, acts = g(latent.cuda(), noise=[n.cuda() for n in noise], truncation=psi.cuda(), input_is_latent=True,
                      randomize_noise=False, return_activation_maps=True)
    # the comparator itself: reference CUDA ops + cuDNN fp32 reproduce the CPU oracle
    e_ref = max([rel_err(a.cpu().numpy(), r.numpy()) for a, r in zip(ref_acts, cpu_acts)] +
                [rel_err(ref_img.cpu().numpy(), cpu_img.numpy())])
    # the product against the reference's CUDA path (fp32 convs)
    e_ours = max([rel_err(a.cpu().numpy(), r.cpu().numpy()) for a, r in zip(acts, ref_acts)] +
                 [rel_err(img.cpu().numpy(), ref_img.cpu().numpy())])
    # for the record: what the reference's DEFAULT GPU setting (TF32 cuDNN convs) does to the same network
    e_tf32 = max([rel_err(a.cpu().numpy(), r.numpy()) for a, r in zip(tf_acts, cpu_acts)] +
                 [rel_err(tf_img.cpu().numpy(), cpu_img.numpy())])
    print(f"reference CUDA path vs CPU oracle {e_ref:.2e}; product vs reference CUDA path {e_ours:.2e}; "
          f"reference with its default TF32 convs vs CPU oracle {e_tf32:.2e}")
    assert e_ref < 1e-4
    assert e_ours < 1e-3
