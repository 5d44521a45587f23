"""Checkpoint wire-format compatibility with the reference — pure host-side code.

The reference writes ``{"model": state_dict, "optimizer": Adam.state_dict(), "ema": {"decay", "shadow", "num_updates"},
"scheduler": LambdaLR.state_dict(), "epoch": N}`` (``Trainer.save_checkpoint`` / ``named_state_dicts``,
ddpm_torch/utils/train.py:264-276), with ``module.`` key prefixes when the model was DDP-wrapped, and reads it back with
``Trainer.load_checkpoint`` (utils/train.py:249-262) and ``generate.py:72-93``.  The native ``UNet`` already has the
reference's ``state_dict`` keys / shapes / order; this module adds the prefix handling, the EMA / optimiser / scheduler
entries (``FusedAdam`` folds the LambdaLR warm-up, so the scheduler entry is synthesised / consumed here) and the
file-name rule, so that checkpoints written by either code base load into the other."""
import re

import torch


def _strip_module_prefix(d):
    """generate.py:83-85 / utils/train.py:256-258: state_dict of a DDP-wrapped model."""
    for k in list(d.keys()):
        if k.startswith("module."):
            d[k.split(".", maxsplit=1)[1]] = d.pop(k)
    return d


def load_weights(model, chkpt, use_ema=False):
    """generate.py:72-93: take ``chkpt["ema"]["shadow"]`` (``use_ema``) or ``chkpt["model"]``; a bare state_dict is accepted
    too ("Try loading checkpoint directly as model weights"); ``module.`` prefixes are stripped.  Raises on key mismatch."""
    if isinstance(chkpt, (str, bytes)) or hasattr(chkpt, "__fspath__"):
        chkpt = torch.load(chkpt, map_location="cpu")
    try:
        sd = chkpt["ema"]["shadow"] if use_ema else chkpt["model"]
    except KeyError:
        sd = chkpt
    sd = _strip_module_prefix(dict(sd))
    model.load_state_dict(sd)
    if hasattr(model, "repack"):
        model.repack()
    return model


def load_checkpoint(chkpt, model, optimizer=None, ema=None, map_location="cpu"):
    """utils/train.py:249-262.  ``optimizer`` is a ``FusedAdam`` (or any object with torch.optim.Adam's state_dict layout),
    ``ema`` an ``EMA``.  Returns the stored epoch (``self.start_epoch = chkpt["epoch"]``)."""
    if isinstance(chkpt, (str, bytes)) or hasattr(chkpt, "__fspath__"):
        chkpt = torch.load(chkpt, map_location=map_location)
    target = getattr(model, "module", model)
    target.load_state_dict(_strip_module_prefix(dict(chkpt["model"])))
    if hasattr(target, "repack"):
        target.repack()
    if optimizer is not None:
        optimizer.load_state_dict(chkpt["optimizer"])
        sch = chkpt.get("scheduler")
        if sch is not None and hasattr(optimizer, "steps"):
            # LambdaLR.last_epoch == number of scheduler steps == number of optimizer steps (utils/train.py:160-163)
            optimizer.steps = int(sch["last_epoch"])
            if "base_lrs" in sch and hasattr(optimizer, "base_lr"):
                optimizer.base_lr = float(sch["base_lrs"][0])
    if ema is not None:
        e = dict(chkpt["ema"])
        e["shadow"] = _strip_module_prefix(dict(e["shadow"]))
        ema.load_state_dict(e)
    return chkpt.get("epoch", 0)


def scheduler_state(optimizer):
    """The ``LambdaLR.state_dict()`` the reference's scheduler would hold after ``optimizer.steps`` steps (train.py:130-132)."""
    return {"base_lrs": [optimizer.base_lr], "last_epoch": optimizer.steps, "_step_count": optimizer.steps + 1,
            "_get_lr_called_within_step": False, "_last_lr": [optimizer.lr], "lr_lambdas": [None]}


def save_checkpoint(chkpt_path, model, optimizer=None, ema=None, **extra_info):
    """utils/train.py:264-272, including the ``_<epoch>.pt`` file-name rule.  Returns the path written."""
    target = getattr(model, "module", model)
    chkpt = [("model", target.state_dict())]
    if optimizer is not None:
        chkpt.append(("optimizer", optimizer.state_dict()))
    if ema is not None:
        chkpt.append(("ema", ema.state_dict()))
    # the reference ALWAYS writes the key (DummyScheduler.state_dict() is None without warm-up) and its loader indexes it
    # unconditionally (utils/train.py:249-262, 264-276)
    chkpt.append(("scheduler", scheduler_state(optimizer) if optimizer is not None and getattr(optimizer, "warmup", 0) > 0 else None))
    for k, v in extra_info.items():
        chkpt.append((k, v))
    if "epoch" in extra_info:
        chkpt_path = re.sub(r"(_\d+)?\.pt", f"_{extra_info['epoch']}.pt", chkpt_path)
    torch.save(dict(chkpt), chkpt_path)
    return chkpt_path
