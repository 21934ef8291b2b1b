import sys, torch
sys.path.insert(0, '.')
import unina_yolo_dla_b200 as uyd
from oracle import custom_graph as cg
from oracle import init as oi
for seed, S in ((3, 640), (1, 640), (1, 320), (3, 320)):
    m = uyd.UninaCustomB200(4, 32).init_synthetic(seed=seed)
    ref = cg.CustomNet(4, 32); ref.load_state_dict(m.state_dict(), strict=True); ref.eval()
    m = m.cuda()
    x = oi.seeded_frames(2, S, seed=11)
    with torch.no_grad():
        want = ref(x)
    got = m(x.cuda()); torch.cuda.synchronize()
    out = []
    for lvl, (pg, pw) in enumerate(zip(got, want)):
        for name, g, w in zip(("cls", "reg"), pg, pw):
            d = (g.cpu() - w).abs(); rng = float(w.abs().max())
            out.append(f"L{lvl}{name} {float(d.max())/rng:.2e} (rms {float(d.pow(2).mean().sqrt())/rng:.1e})")
    print(f"seed {seed} S {S}:", "  ".join(out))
