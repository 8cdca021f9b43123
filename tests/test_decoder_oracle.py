"""Pin the downstream-probe oracle (oracle/mask_decoder_oracle.py) against the reference MaskDecoder_Edge:
golden logits produced by the reference module (oracle/make_golden.py), plus the live module when
/root/reference is present."""
import os

import numpy as np
import pytest
import torch

from oracle import mask_decoder_oracle as mdo
from oracle.ref_import import reference_available

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_decoder_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "decoder_small.npz"))
    got = mdo.decode_mask_logits(mdo.seeded_decoder_state(0), torch.from_numpy(g["emb"]))
    want = torch.from_numpy(g["logits"])
    assert got.shape == want.shape == (1, 1, 64, 64)
    assert float((got - want).abs().max()) < 1e-5
    assert mdo.iou(mdo.binary_mask(got, 256), mdo.binary_mask(want, 256)) == 1.0


def test_iou_helper_and_noise_sensitivity():
    sd = mdo.seeded_decoder_state(0)
    gen = torch.Generator().manual_seed(3)
    emb = torch.randn(1, 256, 16, 16, generator=gen)
    a = mdo.binary_mask(mdo.decode_mask_logits(sd, emb), 256)
    assert mdo.iou(a, a) == 1.0 and 0.05 < float(a.float().mean()) < 0.95      # a non-trivial mask
    b = mdo.binary_mask(mdo.decode_mask_logits(sd, emb + 0.5 * torch.randn(emb.shape, generator=gen)), 256)
    assert mdo.iou(a, b) < 0.99                                                  # the probe does react to errors


@pytest.mark.skipif(not reference_available(), reason="/root/reference not present (GPU box)")
def test_decoder_oracle_matches_live_reference_keys_and_output():
    from oracle.ref_import import load_reference_decoder_modules
    md, tr = load_reference_decoder_modules()
    dec = md.MaskDecoder(num_multimask_outputs=3, transformer_dim=256, iou_head_depth=3, iou_head_hidden_dim=256,
                         transformer=tr.TwoWayTransformer(depth=2, embedding_dim=256, mlp_dim=2048, num_heads=8)).eval()
    shapes = mdo.decoder_param_shapes()
    ref = dec.state_dict()
    assert set(ref) == set(shapes)
    for k, v in ref.items():
        assert tuple(v.shape) == shapes[k], k
    sd = mdo.seeded_decoder_state(7)
    dec.load_state_dict({k: v for k, v in sd.items() if not k.startswith("probe.")}, strict=True)
    emb = torch.randn(1, 256, 8, 8, generator=torch.Generator().manual_seed(7))
    dense = sd["probe.no_mask_embed"].reshape(1, -1, 1, 1).expand(1, -1, 8, 8)
    with torch.no_grad():
        want, _, _ = dec(image_embeddings=emb, interm_embeddings=[], image_pe=mdo.dense_pe(sd, 8),
                         sparse_prompt_embeddings=sd["probe.sparse"], dense_prompt_embeddings=dense,
                         multimask_output=False)
    assert float((mdo.decode_mask_logits(sd, emb) - want).abs().max()) < 1e-5
