"""Checkpoint adapter (landiff_b200/checkpoint.py): a synthetic SAT-layout engine checkpoint written with the
reference's key names round-trips into the drop-in ControlDiffWarp; contract violations are reported."""
import hashlib

import pytest
import torch

from landiff_b200 import checkpoint as C
from landiff_b200.factory import TINY, build_warp


def _engine_checkpoint(tmp_path, warp, extra=None, drop=None):
    g = torch.Generator().manual_seed(7)
    module = {}
    for k, v in warp.state_dict().items():
        if drop and k == drop:
            continue
        module["model." + k] = torch.randn(v.shape, generator=g).to(torch.bfloat16)
    module["model." + C.SEMANTIC_PREFIX + "decoder.conv_in.weight"] = torch.ones(4, 3, 3, 3)
    module["conditioner.embedders.0.transformer.shared.weight"] = torch.zeros(8, 4)
    module["first_stage_model.decoder.conv_out.weight"] = torch.zeros(2, 2)
    module.update(extra or {})
    it = tmp_path / "1000"
    it.mkdir()
    torch.save({"module": module, "iteration": 1000}, it / "mp_rank_00_model_states.pt")
    (tmp_path / "latest").write_text("1000\n")
    return module


def test_engine_checkpoint_round_trip(tmp_path):
    warp = build_warp(TINY)
    module = _engine_checkpoint(tmp_path, warp)
    sem = C.load_engine_checkpoint(warp, str(tmp_path))
    assert list(sem) == ["decoder.conv_in.weight"]
    for k, v in warp.state_dict().items():
        assert torch.equal(v.float(), module["model." + k].to(v.dtype).float()), k
    warp_sd, sem_sd, others = C.split_engine_state(module)
    assert others == ["conditioner", "first_stage_model"] and len(warp_sd) == len(warp.state_dict())


def test_contract_violations_are_reported(tmp_path):
    warp = build_warp(TINY)
    key = next(k for k in warp.state_dict() if k.endswith("query_key_value.weight"))
    (tmp_path / "a").mkdir()
    _engine_checkpoint(tmp_path / "a", warp, drop=key)
    with pytest.raises(KeyError, match="missing"):
        C.load_engine_checkpoint(warp, str(tmp_path / "a"))
    (tmp_path / "b").mkdir()
    _engine_checkpoint(tmp_path / "b", warp, extra={"model." + key: torch.zeros(3, 3)})
    with pytest.raises(ValueError, match="shape mismatch"):
        C.load_engine_checkpoint(warp, str(tmp_path / "b"))
    with pytest.raises(FileNotFoundError):
        C.resolve_checkpoint_path(str(tmp_path / "nope"))


def test_verify_md5(tmp_path):
    (tmp_path / "x.bin").write_bytes(b"landiff")
    good = hashlib.md5(b"landiff").hexdigest()
    (tmp_path / "CHECKSUM.md5").write_text(f"{good}  x.bin\n{'0' * 32}  x.bin.bad\n{good} *missing.bin\n")
    (tmp_path / "x.bin.bad").write_bytes(b"other")
    assert C.verify_md5(str(tmp_path / "CHECKSUM.md5")) == {"x.bin": True, "x.bin.bad": False, "missing.bin": False}


class StubSemanticCond(torch.nn.Module):
    """Shaped like the reference SemanticCond ctor (condition.py:32-45): `dtype` is a REQUIRED keyword-only argument."""

    def __init__(self, width=4, *, dtype):
        super().__init__()
        self.dtype_seen = dtype
        self.conv_out = torch.nn.Conv2d(width, 16, 3, padding=1)

    def forward(self, indexs=None):
        raise NotImplementedError


def _warp_with_conditioner():
    from landiff_b200 import dit
    from landiff_b200.factory import network_params

    pc = network_params(TINY, True)
    pc["modules"]["semantic_condition_config"] = {"target": "test_checkpoint.StubSemanticCond", "params": {"width": 4}}
    ctrl = dit.ControlDiffusionTransformer(**pc, dtype="bf16")
    main = dit.DiffusionTransformer(**network_params(TINY, False), dtype="bf16")
    return dit.ControlDiffWarp(dit.OpenAIWrapper(main), dit.OpenAIWrapper(ctrl), None, True)


def test_semantic_conditioner_gets_dtype_kwarg_like_the_reference():
    """dit_video_concat.py:926-928 passes dtype=self.dtype; a conditioner with a required kw-only dtype must build."""
    warp = _warp_with_conditioner()
    cond = warp.control_model.diffusion_model.semantic_conditioner
    assert isinstance(cond, StubSemanticCond) and cond.dtype_seen == torch.bfloat16
    assert any(k.startswith(C.SEMANTIC_PREFIX) for k in warp.state_dict())


def test_engine_checkpoint_with_parameterised_conditioner(tmp_path):
    """The conditioner's tensors are loaded into the attached submodule and do not count as missing DiT keys."""
    warp = _warp_with_conditioner()
    g = torch.Generator().manual_seed(3)
    module = {"model." + k: torch.randn(v.shape, generator=g) for k, v in warp.state_dict().items()}
    it = tmp_path / "7"
    it.mkdir()
    torch.save({"module": module}, it / "mp_rank_00_model_states.pt")
    (tmp_path / "latest").write_text("7\n")
    sem = C.load_engine_checkpoint(warp, str(tmp_path))
    assert sorted(sem) == ["conv_out.bias", "conv_out.weight"]
    cond = warp.control_model.diffusion_model.semantic_conditioner
    assert torch.equal(cond.conv_out.weight.float(), module["model." + C.SEMANTIC_PREFIX + "conv_out.weight"].to(cond.conv_out.weight.dtype).float())
    # a missing conditioner tensor is reported under strict loading
    del module["model." + C.SEMANTIC_PREFIX + "conv_out.bias"]
    torch.save({"module": module}, it / "mp_rank_00_model_states.pt")
    with pytest.raises(KeyError, match="semantic conditioner"):
        C.load_engine_checkpoint(warp, str(tmp_path))
    C.load_engine_checkpoint(warp, str(tmp_path), strict=False)
