"""CPU-side checks: the C-ABI library builds/loads and exports every symbol include/*.h declares;
entries fail loudly (no fallback) without a GPU; host-side helpers."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from llava_reward_b200.build import build
    build()
    from llava_reward_b200 import _lib
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "llava_reward_b200.h")).read()
    declared = set(re.findall(r"^int (lr_\w+)\(", hdr, flags=re.M))
    assert len(declared) >= 15
    from llava_reward_b200 import _lib
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.lr_version() == 100


def test_argument_validation_without_gpu(lib):
    from llava_reward_b200 import _lib as L
    # bad arguments are rejected before any CUDA call
    assert lib.lr_gemm_bf16(None, 0, None, 0, None, 0, 1, 1, 1, 0, None, None, 0, 0, None) == -1
    assert lib.lr_rmsnorm_bf16(None, 0, None, None, None, 0, 1, 8, 1e-5, None) == -1
    assert lib.lr_attention_bf16(None, None, None, None, 0, 0, 1, 1, None, None, 1, 64, 0, 1.0, 0, None) == -1
    assert lib.lr_softmax_rows_bf16(None, 8, 1, 8, 8, 1.0, None) == -1
    assert lib.lr_masked_mean_rows_bf16(None, 256, None, None, 256, 1, 1, 256, None) == -1
    assert lib.lr_gather_rows_bf16(None, 8, None, None, 8, 1, 8, None) == -1
    # the model surface honours the reference's attributes without a device (rw_model_general_preference.py:327-333)
    from llava_reward_b200.config import RewardConfig as _RC
    from llava_reward_b200.model import B200RewardModel as _M
    from llava_reward_b200.synth import SynthProvider as _SP
    _cfg = _RC(num_layers=1, clip_layers=1)
    _m = _M(_cfg, _SP(_cfg))
    assert _m.layer_id == 32 and _m.mean_hidden_state is None and _m.training is False
    assert _m.train().training is True and _m.eval().training is False and _m.train(False).training is False
    if not torch.cuda.is_available():
        assert lib.lr_device_check() != 0
        from llava_reward_b200 import ops
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            ops.rmsnorm(torch.zeros(1, 8), torch.zeros(8), torch.zeros(1, 8), 1, 8, 1e-5)
        from llava_reward_b200.config import RewardConfig
        from llava_reward_b200.model import B200RewardModel
        from llava_reward_b200.synth import SynthProvider
        cfg = RewardConfig(num_layers=1, clip_layers=1)
        m = B200RewardModel(cfg, SynthProvider(cfg))
        with pytest.raises(RuntimeError):
            m.to("cpu")
        with pytest.raises(RuntimeError):
            m.custom_forward(torch.zeros(1, 4, dtype=torch.long), torch.ones(1, 4, dtype=torch.long),
                             torch.zeros(1, 17, 3, 336, 336), torch.tensor([[336, 336]]))


def test_synth_is_deterministic_and_well_scaled():
    from llava_reward_b200.synth import hash_normal, hash_randint
    a = hash_normal("x.weight", (1000, 64), 0.02, 1234)
    b = hash_normal("x.weight", (1000, 64), 0.02, 1234, chunk=777)
    assert torch.equal(a, b)
    assert abs(a.std().item() - 0.02) < 5e-4 and abs(a.mean().item()) < 5e-4
    assert not torch.equal(a, hash_normal("y.weight", (1000, 64), 0.02, 1234))
    r = hash_randint("t", 1000, 3, 31999, 7)
    assert r.min() >= 3 and r.max() < 31999
    # pinned values: the generator must not drift between environments (CPU here, CUDA on the B200 box)
    v = hash_normal("pin", (4,), 1.0, 1)
    assert torch.allclose(v, torch.tensor(PINNED), atol=0, rtol=0), v.tolist()


PINNED = [-1.5067435503005981, 0.9357186555862427, -0.2206028401851654, -0.5484809875488281]


def test_gate_up_interleave_permutation():
    I = 512
    nb = I // 128
    perm = torch.arange(2 * I).view(2, nb, 128).permute(1, 0, 2).reshape(-1)
    assert perm[:128].tolist() == list(range(128))
    assert perm[128:256].tolist() == list(range(I, I + 128))
    assert perm[256:384].tolist() == list(range(128, 256))
    assert sorted(perm.tolist()) == list(range(2 * I))


def test_num_image_tokens():
    from llava_reward_b200.config import num_image_tokens
    assert num_image_tokens(1344, 1344) == 2509
    assert num_image_tokens(1008, 1344) == 1921
    assert num_image_tokens(336, 672) == 457
