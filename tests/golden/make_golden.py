"""Generates the committed golden fixtures from the REFERENCE's own code and artefact
(run in the build container, where /root/reference exists):

  j_regressor_nnz.npz   the 107 non-zeros of models/retrained_J_Regressor.pt (+ its sha256)
  retrained_J_Regressor.pt   the artefact itself, byte for byte (469 240 B of weights saved from cuda:0,
                        requires_grad, column-major): /root/reference does not exist on the GPU box, and
                        "loads unchanged" can only be shown on the real bytes (bench.py --regressor shipped
                        and the loader tests read it through jrr_b200.load_j_regressor)
  ref_utils_golden.npz  inputs/outputs of the reference's rot6d_to_rotmat, move_pelvis,
                        find_joints (on the oracle SMPL), evaluate, Discriminator.forward,
                        computed by importing /root/reference/scripts/{utils,discriminator}.py

    python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("JRR_REFERENCE_ROOT", "/root/reference")

import jrr_b200 as jrr  # noqa: E402
from oracle import jrr_oracle as O  # noqa: E402


def main():
    art = os.path.join(REF, "models", "retrained_J_Regressor.pt")
    sha = hashlib.sha256(open(art, "rb").read()).hexdigest()
    J = torch.load(art, map_location="cpu", weights_only=True).detach().contiguous()
    import shutil
    shutil.copyfile(art, os.path.join(HERE, "retrained_J_Regressor.pt"))
    r, c = torch.nonzero(J, as_tuple=True)
    np.savez(os.path.join(HERE, "j_regressor_nnz.npz"), row=r.numpy().astype(np.int32),
             col=c.numpy().astype(np.int32), val=J[r, c].numpy(), sha256=np.array(sha))

    u, d, e = O.load_reference_modules(REF)
    model = jrr.synthetic.make_smpl_model(0)
    smpl = O.OracleSMPL(model)
    g = torch.Generator().manual_seed(1234)
    x6 = torch.randn(6, 24, 6, generator=g)
    R = u.rot6d_to_rotmat(x6.reshape(-1, 6)).view(6, 24, 3, 3)
    betas = torch.randn(6, 10, generator=g)
    pred = u.find_joints(smpl, betas, R[:, :1], R[:, 1:], J, mask=u.find_j_reg_mask(J))
    pelvis = u.move_pelvis(pred)
    gt = 1000 * pelvis + 7.0 * torch.randn(6, 17, 3, generator=g)
    mpjpe, pampjpe = u.evaluate(pred, gt)
    torch.manual_seed(0)
    D = d.Discriminator()
    scores = D(x6).detach()
    np.savez(os.path.join(HERE, "ref_utils_golden.npz"), x6=x6.numpy(), rotmat=R.numpy(),
             betas=betas.numpy(), find_joints=pred.detach().numpy(), move_pelvis=pelvis.detach().numpy(),
             gt_mm=gt.detach().numpy(), mpjpe=np.float64(mpjpe), pa_mpjpe=np.float64(pampjpe),
             critic_scores=scores.numpy(), mask_sum=np.float64(u.find_j_reg_mask(J).sum().item()))
    # Shape_Discriminator (scripts/discriminator.py:57-74): default init under seed 0, scores on `betas`
    torch.manual_seed(0)
    S = d.Shape_Discriminator()
    np.savez(os.path.join(HERE, "ref_shape_critic_golden.npz"), betas=betas.numpy(),
             shape_scores=S(betas).detach().numpy(),
             **{k.replace(".", "__"): v.detach().numpy() for k, v in S.state_dict().items()})
    # critic training step, optimize.py:113-123,276-293, run with the reference's own modules:
    # two Adam(lr=1e-3) steps of MSE(D(fake),0)+MSE(D(real),1) for both discriminators
    import torch.nn as nn
    torch.manual_seed(0)
    D = d.Discriminator()
    torch.manual_seed(0)
    S = d.Shape_Discriminator()
    optD = torch.optim.Adam(D.parameters(), lr=1e-3); optS = torch.optim.Adam(S.parameters(), lr=1e-3)
    x6_real = x6 + 0.3 * torch.randn(6, 24, 6, generator=g)
    betas_real = betas + 0.5 * torch.randn(6, 10, generator=g)
    mse = nn.MSELoss()
    losses, losses_s = [], []
    for _ in range(2):
        pf, pr = D(x6), D(x6_real)
        l = mse(pf, torch.zeros(pf.shape)) + mse(pr, torch.ones(pf.shape))
        optD.zero_grad(); l.backward(); optD.step(); losses.append(l.item())
        pf, pr = S(betas), S(betas_real)
        l = mse(pf, torch.zeros(pf.shape)) + mse(pr, torch.ones(pf.shape))
        optS.zero_grad(); l.backward(); optS.step(); losses_s.append(l.item())
    sdD = D.state_dict()
    np.savez(os.path.join(HERE, "ref_critic_train_golden.npz"), x6_fake=x6.numpy(), x6_real=x6_real.numpy(),
             betas_fake=betas.numpy(), betas_real=betas_real.numpy(), losses=np.array(losses), losses_shape=np.array(losses_s),
             conv0_w=sdD["conv_operations.0.weight"].numpy(), conv2_b=sdD["conv_operations.2.bias"].numpy(),
             lin3_w=sdD["linears.3.weight"].numpy(), lin3_b=sdD["linears.3.bias"].numpy(),
             w1_block=sdD["linear_operations.0.weight"][:4, :8].numpy(), b2=sdD["linear_operations.2.bias"].numpy(),
             w3=sdD["linear_operations.4.weight"].numpy(),
             **{"shape__" + k.replace(".", "__"): v.detach().numpy() for k, v in S.state_dict().items()})
    # crop arithmetic of the data loader (scripts/data.py:123-138,216-270) from the reference's own find_crop
    rd = O.load_reference_data_module(REF)
    gb = torch.Generator().manual_seed(77)
    lo = 100 + 300 * torch.rand(5, 2, generator=gb)
    hi = lo + 150 + 400 * torch.rand(5, 2, generator=gb)
    bboxes = torch.stack([lo[:, 0], lo[:, 1], hi[:, 0], hi[:, 1]], dim=1)          # [min_y, min_x, max_y, max_x]
    intr = torch.zeros(5, 3, 3)
    intr[:, 0, 0] = 1145 + torch.rand(5, generator=gb); intr[:, 1, 1] = 1144 + torch.rand(5, generator=gb)
    intr[:, 0, 2] = 500 + 20 * torch.rand(5, generator=gb); intr[:, 1, 2] = 510 + 20 * torch.rand(5, generator=gb)
    intr[:, 2, 2] = 1
    _, mnx, mny, sc, intr_out = rd.find_crop(torch.zeros(5, 3, 64, 64), bboxes, intr)
    j2d = 1000 * torch.rand(5, 17, 2, generator=gb)
    rep = j2d.clone()                                   # data.py:134-138 on the reference's crop parameters
    rep[..., 0] -= mnx[:, None]; rep[..., 1] -= mny[:, None]
    rep /= sc[:, None, None]; rep /= 1000 / 224
    np.savez(os.path.join(HERE, "ref_data_crop_golden.npz"), bboxes=bboxes.numpy(), intrinsics=intr.numpy(),
             min_x=mnx.numpy(), min_y=mny.numpy(), scale=sc.numpy(), intrinsics_out=intr_out.numpy(),
             gt_j2d=j2d.numpy(), gt_j2d_repositioned=rep.numpy())
    print("wrote fixtures; artefact sha256", sha)


if __name__ == "__main__":
    main()
