import os, sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import nerf_for_angiography_b200 as A
g = np.load("/root/repo/tests/golden/cppn_barf_4x128.npz")
for prec in ("fp32", "bf16"):
    d = {'num_early_layers': 4, 'num_late_layers': 0, 'num_filters': 128, 'num_input_channels': 3, 'num_output_channels': 1,
         'num_input_channels_views': 0, 'use_bias': True, 'pos_enc': 'barf', 'pos_enc_basis': 5, 'act_func': 'relu',
         'fourier_sigma': 5, 'num_img': 1, 'device': torch.device("cuda"), 'precision': prec}
    model = A.CPPN(d)
    sd = {k[3:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("sd:")}
    model.load_state_dict({**sd, "barf_weights": torch.zeros(15)}); model = model.to("cuda")
    x = torch.from_numpy(g["x"]).cuda()
    for ai, a in enumerate(g["alphas"]):
        model.update_barf_alpha(float(a), 'pts'); model.zero_grad()
        y = model(x); yref = torch.from_numpy(g[f"a{ai}_y"])
        (y * torch.from_numpy(g[f"a{ai}_gout"]).cuda()).sum().backward()
        errs = []
        for k in g.files:
            if k.startswith(f"a{ai}_grad:"):
                name = k.split(":", 1)[1]
                got = dict(model.named_parameters())[name].grad.cpu(); ref = torch.from_numpy(g[k])
                errs.append(f"{name.replace('early_pts_layers.','L').replace('output_linear.0','out')}={float((got-ref).norm()/max(float(ref.norm()),1e-9)):.3f}")
        print(prec, f"alpha={a}: y err {float((y.detach().cpu()-yref).abs().max()):.4f} (|y|max {float(yref.abs().max()):.2f})", " ".join(errs))
