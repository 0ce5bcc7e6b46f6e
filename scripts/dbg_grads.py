import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import test_model_parity as T
w, om, pm, ob, pb = T.build("toyotagraph", "tiny", 6, 12)
om.train(); pm.train(); pm.pos_embed.p = 0.0
for m in (om, pm):
    m.poi_distance_model.eval(); m.poi_cat_model.eval()
lref = om.training_loss(ob); lref.backward()
lgot = pm.training_step(pb); lgot.backward()
print("loss", lref.item(), lgot.item())
ref_g = {k: p.grad for k, p in om.named_parameters() if p.grad is not None}
rows = []
for k, p in pm.named_parameters():
    if k in ref_g and p.grad is not None:
        r = ref_g[k]; g = p.grad.float().cpu()
        sc = r.abs().max().item()
        if sc > 0:
            cos = torch.nn.functional.cosine_similarity(g.flatten(), r.flatten(), dim=0).item()
            rows.append((k, (g - r).abs().max().item() / sc, sc, cos, g.norm().item() / (r.norm().item() + 1e-30)))
for r in rows:
    print("%-55s relerr %.4f  max %.3e  cos %.4f  normratio %.4f" % r)
