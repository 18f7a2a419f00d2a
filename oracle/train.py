"""Oracle training step: loss and parameter gradients by CPU autograd over the restated model maths.

TEST INFRASTRUCTURE ONLY — see oracle/__init__.py.  Restates, with autograd enabled, exactly what
Trainer._train_epoch's inner body computes for one batch (reference src/trainer/trainer.py:250-253: forward, DiceCE,
backward) for UNet3D (src/models/backbones/unet.py:165-200) and DualEncoder (dual_encoder.py:112-199, fusion
concat | add | mean | attention gate; model.backbone.norm = instance | group | batch | none), dropout off.  state_dict keys WITHOUT the "backbone." prefix.
"""
import torch
import torch.nn.functional as F2


def train_step(kind, sd, cfgkw, x, y, dtype=torch.float64, dice_weight=0.5, ce_weight=0.5, input_grad=False):
    """Oracle training step on the CPU: loss + parameter gradients by autograd over the oracle restatement
    (input_grad: the gradient w.r.t. x is returned under the key "__input__")."""
    params = {k: v.detach().to("cpu", dtype).requires_grad_(True) for k, v in sd.items() if v.is_floating_point()}
    xx = x.detach().to("cpu", dtype).requires_grad_(bool(input_grad))
    # the oracle forwards detach their parameters; re-state them here with autograd enabled
    norm = cfgkw.get("norm", "instance")     # model.backbone.norm — reference unet.py:29-41 (GroupNorm(8, C) | BatchNorm3d)

    def nrm(prefix, t):
        if norm == "instance":
            return F2.instance_norm(t, eps=1e-5)
        if norm == "group":
            return F2.group_norm(t, 8, params[prefix + ".weight"], params[prefix + ".bias"], eps=1e-5)
        if norm == "batch":                  # train mode: batch statistics
            return F2.batch_norm(t, None, None, params[prefix + ".weight"], params[prefix + ".bias"], True, 0.1, 1e-5)
        return t                             # anything else: nn.Identity

    def block(prefix, t):
        for i in (1, 2):
            t = F2.conv3d(t, params[f"{prefix}.conv{i}.weight"], params[f"{prefix}.conv{i}.bias"], padding=1)
            t = F2.relu(nrm(f"{prefix}.norm{i}", t))
        return t
    def up(prefix, t, skip):
        t = F2.conv_transpose3d(t, params[f"{prefix}.up.weight"], params[f"{prefix}.up.bias"], stride=2)
        return block(f"{prefix}.conv", torch.cat([t, skip], 1))
    if kind == "unet":
        L = cfgkw["L"]
        t = block("init_conv", xx)
        feats = [t]
        for i in range(L - 1):
            t = block(f"encoders.{i}.conv", F2.max_pool3d(t, 2))
            feats.append(t)
        dec = "decoders"
    else:
        L, M, fusion = cfgkw["L"], cfgkw["M"], cfgkw["fusion"]
        allf = []
        for m in range(M):
            t = block(f"encoders.{m}.init_conv", xx[:, m:m + 1])
            fl = [t]
            for i in range(L - 1):
                t = block(f"encoders.{m}.blocks.{i}.conv", F2.max_pool3d(t, 2))
                fl.append(t)
            allf.append(fl)
        feats = []
        for l in range(L):
            lf = [allf[m][l] for m in range(M)]
            if fusion == "concat":
                feats.append(F2.conv3d(torch.cat(lf, 1), params[f"fusion_proj.{l}.weight"], params[f"fusion_proj.{l}.bias"]))
            elif fusion == "add":
                feats.append(sum(lf))
            elif fusion == "attention":  # CrossModalAttention, dual_encoder.py:243-254
                pooled = torch.cat([t_.mean(dim=(2, 3, 4)) for t_ in lf], dim=1)
                hdn = F2.relu(F2.linear(pooled, params[f"fusion_layers.{l}.attention.2.weight"], params[f"fusion_layers.{l}.attention.2.bias"]))
                wts = torch.softmax(F2.linear(hdn, params[f"fusion_layers.{l}.attention.4.weight"], params[f"fusion_layers.{l}.attention.4.bias"]), dim=1)
                feats.append(sum(wts[:, m_].view(-1, 1, 1, 1, 1) * lf[m_] for m_ in range(M)))
            else:
                feats.append(torch.stack(lf).mean(0))
        t = feats[-1]
        dec = "decoder"
    for j, skip in enumerate(reversed(feats[:-1])):
        t = up(f"{dec}.{j}", t, skip)
    if cfgkw.get("drop") is not None:        # Dropout3d before out_conv (unet.py:162, 198 / dual_encoder.py:201): the
        t = t * cfgkw["drop"].detach().to("cpu", dtype)[:, :, None, None, None]   # caller's [n, C] mask / (1 - p)
    logits = F2.conv3d(t, params["out_conv.weight"], params["out_conv.bias"])
    p = torch.softmax(logits, 1)
    C = p.shape[1]
    tt = F2.one_hot(y.cpu().long(), C).movedim(-1, 1).to(dtype)
    I, U = (p * tt).flatten(2).sum(-1), p.flatten(2).sum(-1) + tt.flatten(2).sum(-1)
    loss = dice_weight * (1 - (2 * I + 1) / (U + 1)).mean() + ce_weight * F2.cross_entropy(logits, y.cpu().long())
    loss.backward()
    grads = {k: v.grad for k, v in params.items()}
    if input_grad:
        grads["__input__"] = xx.grad
    return loss.item(), grads, logits.detach()


