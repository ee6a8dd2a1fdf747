import sys, torch
sys.path.insert(0, '.')
import vae_gan_b200 as v
from oracle import vaegan_oracle as O
from tests.gpu_util import *
g = torch.load('tests/golden/discriminator_fwd_bwd.pt')
sp = dict(g["spec"]); spec = O.DiscriminatorSpec(**sp)
P = O.make_discriminator_params(spec, seed=g["seed_d"]); P.update({k: t.clone() for k, t in g["params"].items()})
with v.compute_dtype(torch.float32):
    D = v.Discriminator(v.ResBlockDiscriminator, sp["num_stride_conv1"], sp["num_features_conv1"], list(sp["num_blocks"]), list(sp["num_strides_res"]), list(sp["num_features_res"]), input_size=sp["input_size"])
    load_params_into(D, P); D = D.to(dev()).train()
    v.rng.seed = 0x5EED5EED; v.rng.reset_sites()
    x = g["x"].to(dev()).requires_grad_(True)
    logits = D(x)
    (logits * g["logit_weights"].to(dev())).sum().backward()
    for k, p in D.named_parameters():
        w = g["grads"][k]
        if isinstance(w, dict): continue
        e = relmax(p.grad, w)
        if e > 1e-4:
            print(k, 'err', e, 'ours', p.grad.flatten()[:6].tolist(), 'want', w.flatten()[:6].tolist(), 'maxabs', float(w.abs().max()))
