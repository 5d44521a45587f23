"""Checkpoint wire format (ddpm_torch/utils/train.py:249-276, generate.py:72-93) against a checkpoint written by the
UNMODIFIED reference UNet / EMA + torch Adam / LambdaLR after 3 optimisation steps with DDP-style ``module.`` prefixes
(tests/golden/checkpoint_micro.pt, oracle/gen_golden.py::gen_checkpoint)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _model(cfg):
    import ddpm_torch_b200 as D
    return D.UNet(**{k: cfg[k] for k in ("in_channels", "hid_channels", "out_channels", "ch_multipliers", "num_res_blocks", "apply_attn")},
                  drop_rate=cfg["drop_rate"])


def test_load_weights_model_and_ema_with_module_prefix(golden):
    from ddpm_torch_b200 import checkpoint as K
    fx = golden("checkpoint_micro.pt")
    chk = fx["chkpt"]
    assert all(k.startswith("module.") for k in chk["model"]) and set(chk) == {"model", "optimizer", "ema", "scheduler", "epoch"}
    for use_ema, src in ((False, chk["model"]), (True, chk["ema"]["shadow"])):
        m = _model(fx["cfg"])
        K.load_weights(m, chk, use_ema=use_ema)
        assert [k for k, _ in m.named_parameters()] == [k[len("module."):] for k in chk["model"]]      # same keys, same order
        for k, p in m.named_parameters():
            assert torch.equal(p.detach(), src["module." + k]), k
    # a bare state_dict is accepted (generate.py:80-81), a foreign one is rejected
    m = _model(fx["cfg"])
    K.load_weights(m, {k[len("module."):]: v for k, v in chk["model"].items()})
    with pytest.raises(RuntimeError):
        K.load_weights(_model(fx["cfg"]), {"nope": torch.zeros(1)})
    # the golden dict was not mutated by the prefix stripping
    assert all(k.startswith("module.") for k in chk["model"])


@pytest.mark.gpu
def test_checkpoint_round_trip_with_reference_format(golden, tmp_path):
    from ddpm_torch_b200 import checkpoint as K
    from ddpm_torch_b200.optim import EMA, FusedAdam
    fx = golden("checkpoint_micro.pt")
    chk = fx["chkpt"]
    m = _model(fx["cfg"]).cuda()
    ema = EMA(m, decay=0.5)
    opt = FusedAdam(m, lr=1e-3, warmup=5, grad_norm=1.0, ema=ema)
    epoch = K.load_checkpoint(chk, m, optimizer=opt, ema=ema)
    assert epoch == 7
    names = [k for k, _ in m.named_parameters()]
    for i, k in enumerate(names):
        assert torch.equal(dict(m.named_parameters())[k].detach().cpu(), chk["model"]["module." + k])
        assert torch.equal(ema.shadow[k].cpu(), chk["ema"]["shadow"]["module." + k])
        st = chk["optimizer"]["state"][i]
        off = m._meta[i][2]; n = st["exp_avg"].numel()
        assert torch.equal(opt.exp_avg[off:off + n].cpu(), st["exp_avg"].reshape(-1))
        assert torch.equal(opt.exp_avg_sq[off:off + n].cpu(), st["exp_avg_sq"].reshape(-1))
    assert opt.steps == 3 and ema.num_updates == chk["ema"]["num_updates"] == 2 and ema.decay == chk["ema"]["decay"]
    assert opt.base_lr == 2e-4 and opt.lr == pytest.approx(chk["scheduler"]["_last_lr"][0], rel=1e-12)      # next step's lr
    # save in the reference format, then load the file into STOCK torch objects exactly as utils/train.py:249-262 would
    path = K.save_checkpoint(str(tmp_path / "ddpm_micro.pt"), m, optimizer=opt, ema=ema, epoch=8)
    assert path.endswith("ddpm_micro_8.pt")
    back = torch.load(path, map_location="cpu")
    assert set(back) == {"model", "optimizer", "ema", "scheduler", "epoch"} and back["epoch"] == 8
    ref_params = [torch.nn.Parameter(v.clone()) for v in chk["model"].values()]
    stock = torch.optim.Adam(ref_params, lr=2e-4)
    sch = torch.optim.lr_scheduler.LambdaLR(stock, lr_lambda=lambda t: min((t + 1) / 5, 1.0))
    stock.load_state_dict(back["optimizer"]); sch.load_state_dict(back["scheduler"])
    assert sch.last_epoch == 3 and stock.param_groups[0]["lr"] == pytest.approx(opt.lr, rel=1e-12)
    for i, p in enumerate(ref_params):
        assert float(stock.state[p]["step"]) == 3.0
        assert torch.equal(stock.state[p]["exp_avg"], chk["optimizer"]["state"][i]["exp_avg"])
    assert list(back["model"].keys()) == names and set(back["ema"]) == {"decay", "shadow", "num_updates"}
    for k in names:
        assert torch.equal(back["model"][k].cpu(), chk["model"]["module." + k])
        assert torch.equal(back["ema"]["shadow"][k].cpu(), chk["ema"]["shadow"]["module." + k])
    # and the weights drive the engine: EMA context swaps them in and out (utils/train.py:179)
    x = torch.randn(2, 3, 16, 16, device="cuda"); t = torch.randint(1000, (2,), device="cuda")
    with torch.no_grad():
        y0 = m.eval()(x, t)
        with ema:
            y1 = m(x, t)
        y2 = m(x, t)
    assert torch.isfinite(y0).all() and (y0 - y1).abs().max() > 0 and torch.allclose(y0, y2, atol=2e-2, rtol=2e-2)
