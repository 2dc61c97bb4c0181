"""CPU-side checks of the ViT drop-in: checkpoint contract and host logic (no CUDA needed)."""
import numpy as np
import pytest
import torch

from rgb_no_more_b200 import vit as V
from rgb_no_more_b200 import train_step as TS
from tests.helpers import load


def _vitti():
    return V.ViT(patch_size=16, emb_size=192, depth=12, n_classes=1000, drop_p=0.0, num_heads=3, head_size=64,
                 pixel_space="DCT", ver=1, use_subblock=True)


def test_state_dict_keys_and_shapes_match_reference():
    g = load("embed_vit.npz")
    sd = _vitti().state_dict()
    assert sorted(sd.keys()) == [str(k) for k in g["state_keys"]]
    shapes = [",".join(map(str, sd[k].shape)) for k in sorted(sd.keys())]
    assert shapes == [str(s) for s in g["state_shapes"]]
    assert len(sd) == 152


def test_reference_kwargs_surface():
    # utils.get_model passes exactly these (pipeline_utils.py:335-349)
    m = V.ViT(in_channels=3, patch_size=16, emb_size=384, depth=12, n_classes=1000, drop_p=0.0, device="cpu",
              dtype=torch.float32, num_heads=6, head_size=64, pixel_space="dct", ver=1, use_subblock=True)
    assert sum(p.numel() for p in m.parameters()) == 21975016       # SURVEY.md 2.2 [probed]
    assert sum(p.numel() for p in _vitti().parameters()) == 5642728
    with pytest.raises(NotImplementedError):
        V.ViT(pixel_space="RGB")
    with pytest.raises(NotImplementedError):
        V.ViT(pixel_space="DCT", ver=2, emb_size=192, num_heads=3)
    with pytest.raises(NotImplementedError):
        V.ViT(pixel_space="DCT", ver=1, emb_size=192, num_heads=3, drop_p=0.1)


def test_no_cpu_fallback():
    m = _vitti()
    with pytest.raises(Exception):
        m(torch.zeros(1, 196, 384))


def test_posemb_matches_reference_formula():
    g = load("embed_vit.npz")
    # tokens = Linear(embed_in) + posemb -> posemb = tokens - Linear(embed_in) for the golden weights
    from tests.helpers import seeded_state_dict
    m = _vitti()
    m.load_state_dict(seeded_state_dict(m))
    w, b = m.patchembed.projection[0].weight, m.patchembed.projection[0].bias
    lin = torch.nn.functional.linear(torch.from_numpy(g["embed_in"]).reshape(196, 384), w, b)
    pe = V.sincos_posemb(14, 14, 192, "cpu")
    assert float((lin + pe - torch.from_numpy(g["tokens"])).abs().max()) < 2e-5


def test_embed_input_compat_path_matches_reference():
    g = load("embed_vit.npz")
    from tests.helpers import golden_vit_inputs
    yf, cf = golden_vit_inputs(g["input_seed"])
    e = V.embed_input_from_planes(yf, cf)[0].reshape(14, 14, 384).numpy()
    assert np.array_equal(e[..., 256:], g["embed_in"][..., 256:])
    assert np.abs(e[..., :256] - g["embed_in"][..., :256]).max() < 1e-5


def test_lr_schedule():
    st = TS.TrainStage.__new__(TS.TrainStage)
    st.base_lr, st.warmup_steps, st.total_steps = 3e-3, 10, 110
    # reference indexing (train.py:149-152): step s runs with current_itr = s + 1 -> LR * (s + 2) / WARMUP during warm-up,
    # cosine from current_itr = WARMUP on (full trace vs a real torch scheduler: tests/test_train_schedule_cpu.py)
    assert abs(st._lr(0) - 6e-4) < 1e-12 and abs(st._lr(8) - 3e-3) < 1e-12
    assert abs(st._lr(9) - 3e-3) < 1e-12 and abs(st._lr(59) - 1.5e-3) < 1e-9 and st._lr(109) < 1e-12
