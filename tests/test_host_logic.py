"""Host-side logic that needs no GPU: weight naming / packing, schedule, pipeline argument handling."""
import os

import pytest
import torch
import torch.nn.functional as F

from diffsensei_b200.config import RESAMPLER_TINY, SDXL_MANGA, TINY
from diffsensei_b200.scheduler import DDIMScheduler
from diffsensei_b200.weights import (pack_conv3x3, pack_geglu, random_state_dict, resampler_param_shapes,
                                     unet_param_shapes)
from oracle.ddim import DDIMSchedule
from oracle.resampler import OracleResampler
from oracle.unet import OracleUNet


def test_unet_key_names_and_shapes_match_oracle_module():
    sd = OracleUNet(TINY).state_dict()
    sh = unet_param_shapes(TINY)
    assert set(sd) == set(sh)
    assert all(tuple(sd[k].shape) == sh[k] for k in sh)
    assert "dialog_bbox_embedding" in sh
    assert "down_blocks.1.attentions.0.transformer_blocks.0.attn2.processor.to_k_ip.weight" in sh


def test_sdxl_topology_counts():
    sh = unet_param_shapes(SDXL_MANGA)
    n = sum(torch.Size(s).numel() for s in sh.values())
    assert abs(n - 2.908e9) < 5e6                                   # 2.57 B UNet + 0.34 B IP projections
    assert sum(k.endswith("to_k_ip.weight") for k in sh) == 70      # 70 cross-attention sites
    assert sum(".resnets." in k and k.endswith("conv1.weight") for k in sh) == 17


def test_resampler_key_names_match_oracle_module():
    import dataclasses
    kw = dataclasses.asdict(RESAMPLER_TINY)
    sd = OracleResampler(**kw).state_dict()
    sh = resampler_param_shapes(RESAMPLER_TINY)
    assert set(sd) == set(sh) and all(tuple(sd[k].shape) == sh[k] for k in sh)


def test_random_state_dict_clones_ip_projections():
    sd = random_state_dict(unet_param_shapes(TINY), 0, "cpu")
    k = "mid_block.attentions.0.transformer_blocks.0.attn2"
    assert torch.equal(sd[k + ".processor.to_k_ip.weight"], sd[k + ".to_k.weight"])     # unet.py:72-75


def test_pack_geglu_is_a_row_permutation_with_the_documented_block_structure():
    c = 64
    w, b = (torch.randn(8 * c, c) / 8).to(torch.bfloat16).float(), torch.randn(8 * c)
    wp, bp = pack_geglu(w, b)
    x = torch.randn(5, c)
    val, gate = F.linear(x, w, b).chunk(2, dim=-1)
    want = val * F.gelu(gate)
    y = F.linear(x, wp.float(), bp)                                  # [5, 8c] in packed order
    y = y.reshape(5, -1, 2, 128)
    got = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(5, -1)
    assert torch.allclose(got, want, atol=1e-5, rtol=1e-5)           # pure permutation: exact up to fp32 order


def test_pack_conv3x3_tap_major():
    w = torch.randn(8, 64, 3, 3)
    p = pack_conv3x3(w)
    assert p.shape == (8, 3, 3, 64) and p.dtype == torch.bfloat16
    assert torch.equal(p[3, 1, 2].float(), w[3, :, 1, 2].to(torch.bfloat16).float())


def test_ddim_schedule_matches_oracle():
    a, b = DDIMScheduler(), DDIMSchedule()
    for n in (4, 20, 30, 50):
        assert a.set_timesteps(n) == b.set_timesteps(n)
        assert a.timesteps[0] == (n - 1) * (1000 // n) + 1 and a.timesteps[-1] == 1   # leading spacing, offset 1
        for t in a.timesteps:
            assert a.coefficients(t) == b.coefficients(t)
    assert a.set_timesteps(50)[:3] == [981, 961, 941]


def test_pipeline_check_inputs_raises_like_reference():
    from diffsensei_b200.pipeline import DiffSenseiPipeline
    pipe = DiffSenseiPipeline.__new__(DiffSenseiPipeline)
    with pytest.raises(ValueError, match="`prompt` has to be of type `str`"):
        pipe.check_inputs(None, None, [], None, [])
    with pytest.raises(ValueError, match="can not be input together"):
        pipe.check_inputs("a", None, [object()], torch.zeros(1, 16, 8), [[0, 0, 1, 1]])
    with pytest.raises(ValueError, match="must have the same length as `ip_bbox`"):
        pipe.check_inputs("a", None, [object(), object()], None, [[0, 0, 1, 1]])
    pipe.check_inputs("a", None, [object()], None, [[0, 0, 1, 1]])


def test_fold_layernorm_is_algebraically_layernorm_then_linear():
    """weights.fold_layernorm + the ds_gemm_bf16 "consumer" epilogue formula (include/dsengine.h) restated in fp64:
    rstd * (x W'^T - mean * colsum) + b'  ==  LayerNorm(x) W^T + b."""
    import torch
    from diffsensei_b200.weights import fold_layernorm
    g = torch.Generator().manual_seed(0)
    x = torch.randn(37, 96, generator=g, dtype=torch.float64) * 3 + 1.5
    w = torch.randn(50, 96, generator=g, dtype=torch.float64) / 10
    b = torch.randn(50, generator=g, dtype=torch.float64)
    gamma = 1 + 0.2 * torch.randn(96, generator=g, dtype=torch.float64)
    beta = 0.3 * torch.randn(96, generator=g, dtype=torch.float64)
    want = torch.nn.functional.layer_norm(x, (96,), gamma, beta, 1e-5) @ w.T + b
    w2, b2 = fold_layernorm(w, b, gamma, beta)
    w2, b2 = w2.double(), b2.double()
    s, q = x.sum(1), (x * x).sum(1)
    mean = s / 96
    rstd = (q / 96 - mean * mean + 1e-5).rsqrt()
    got = rstd[:, None] * (x @ w2.T - mean[:, None] * w2.sum(1)[None, :]) + b2
    assert (got - want).abs().max() < 1e-5      # fold_layernorm computes in fp32
    w3, b3 = fold_layernorm(w, None, gamma, beta)
    assert torch.allclose(b3.double(), w @ beta, atol=1e-5)


def test_bench_reference_arm_prints_exactly_one_json_line_on_stdout():
    """bench.py --impl reference (tiny plumbing mode): stdout carries ONE JSON line with the contract's keys; library
    chatter (NCCL prints its version on fd 1) is kept off stdout by bench.main's fd juggling."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    env = dict(os.environ, DS_BENCH_TINY="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_host_thread_probe_returns_a_usable_count():
    import bench
    n = bench.pick_host_threads()
    assert 1 <= n <= (os.cpu_count() or 1)


# ------------------------------------------------------------------------------------------ round-2 additions
def test_oracle_topology_is_its_own_and_agrees_with_the_engine_config():
    """oracle and engine must not share one topology dataclass.  oracle/config.py restates the published
    SDXL + manga values independently; the engine's config has to agree with it field by field."""
    import dataclasses
    from oracle.config import SDXL, TINY as OTINY, OracleUNetConfig
    for mine, theirs in ((SDXL_MANGA, SDXL), (TINY, OTINY)):
        for f in dataclasses.fields(OracleUNetConfig):
            assert tuple(getattr(mine, f.name)) == tuple(getattr(theirs, f.name)) if isinstance(
                getattr(theirs, f.name), tuple) else getattr(mine, f.name) == getattr(theirs, f.name), f.name
        assert mine.time_embed_dim == theirs.time_embed_dim and mine.num_ip_tokens == theirs.num_ip_tokens
    assert OracleUNetConfig.from_any(SDXL_MANGA) == SDXL


def test_oracle_and_reference_arm_never_load_the_product(tmp_path):
    """`import oracle...` and `bench.py --impl reference` must not import diffsensei_b200 (which dlopens
    libdsengine.so): the reference arm's evidence has to be clean of the product's native code."""
    import json
    import subprocess
    import sys
    from conftest import ROOT
    code = ("import sys; import oracle.unet, oracle.ddim, oracle.attention, oracle.resampler, oracle.config; "
            "assert 'diffsensei_b200' not in sys.modules; "
            "assert 'libdsengine' not in open('/proc/self/maps').read(); print('clean')")
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0 and "clean" in r.stdout, r.stderr[-2000:]
    env = dict(os.environ, DS_BENCH_TINY="1")
    r = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "2", "--warmup", "1"], cwd=ROOT,
                       capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["product_package_imported"] is False
    assert line["steps"] == 2 and line["warmup"] == 1 and line["cpu_baseline"]["kind"] == "port"
    assert abs(line["value"] * 8 * line["ms_per_step"] / 1e3 - 1.0) < 0.01          # value = 1 / (8 rows x s/row)
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and "cfg1_measured" in line


def test_analytic_flop_model_matches_survey_numbers():
    from oracle.config import SDXL, unet_flops
    assert abs(unet_flops(SDXL, 8, 128, 128) / 1e12 - 54.8) < 0.05          # cfg2, SURVEY §8d
    assert abs(unet_flops(SDXL, 2, 64, 64) / 1e12 - 3.30) < 0.01            # cfg1
    assert abs(unet_flops(SDXL, 2, 256, 128) / 1e12 - 30.2) < 0.05          # cfg5
    assert abs((unet_flops(SDXL, 8, 128, 128) - unet_flops(SDXL, 8, 128, 128, hoist_kv=True)) / 1e12 - 0.856) < 0.005


def test_pack_cache_is_keyed_on_identity_not_address():
    from diffsensei_b200.attention_processor import _PackCache
    c, calls = _PackCache(), []
    a = torch.zeros(4)
    build = lambda: calls.append(1) or len(calls)
    assert c.get((a,), build) == 1 and c.get((a,), build) == 1
    a.add_(1)                                           # version bump
    assert c.get((a,), build) == 2
    b = torch.zeros(4)                                  # different object (whatever its address)
    assert c.get((b,), build) == 3 and c.get((b,), build) == 3
