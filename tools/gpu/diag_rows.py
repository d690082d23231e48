"""Diagnostic: per-patch (camera, ISO) rows through the three gradient paths (device graph / device plain / host)."""
import copy, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
import numpy as np, torch
from common import synth_batch, make_oracle
from noise_flow_b200 import NoiseFlow, hps_loader, load_checkpoint
from noise_flow_b200.train import DeviceTrainer, loss_and_grad
g = os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "NoiseFlow")
hps = hps_loader(os.path.join(g, "hps.txt")); ck = load_checkpoint(os.path.join(g, "ckpt", "model.ckpt.best"))
x, y = synth_batch(6, cam=2, iso=100, seed=97)
cams = [2.0, 2.0, 0.0, 4.0, 1.0, 2.0]; isos = [100.0, 800.0, 400.0, 100.0, 3200.0, 1600.0]

def rel(a, b):
    worst = (0, "")
    for k in b:
        sc = max(np.abs(b[k]).max(), 0.5)
        e = np.abs(a[k] - b[k]).max() / sc
        if e > worst[0]: worst = (e, k)
    return worst

for training in (False, True):
    res = {}
    for name, graph in (("graph", True), ("plain", False), ("graph2", True), ("plain2", False)):
        nf = NoiseFlow([32, 32, 4], training, copy.copy(hps), variables=ck, device="cuda:0", first_call="inverse")
        tr = DeviceTrainer(nf, max_batch=8, cuda_graph=graph)
        tr.loss_and_grad(x, y, iso=isos, cam=cams, is_training=training)
        res[name] = (tr.loss()[0], tr.gradients())
        if name == "graph":      # replay once more on the same trainer
            tr.loss_and_grad(x, y, iso=isos, cam=cams, is_training=training)
            res["graph_replay"] = (tr.loss()[0], tr.gradients())
    for rep in range(2):
        nf2 = NoiseFlow([32, 32, 4], training, copy.copy(hps), variables=ck, device="cuda:0", first_call="inverse")
        l, _, gr = loss_and_grad(nf2, x, y, iso=isos, cam=cams, is_training=training)
        res["host%d" % rep] = (l, gr)
    print("is_training", training)
    base = res["plain"]
    for k, (l, gr) in res.items():
        print("  %-13s loss %.6f  worst rel diff vs plain %.3e (%s)" % (k, l / 4096, *rel(gr, base[1])))
    if not training:    # oracle: mean of independent per-patch gradients
        acc = None
        for i in range(6):
            orc = make_oracle(hps, ck)
            orc._loss(x[:1], y[:1], iso=[isos[i]], cam=[cams[i]])
            params = {k: v for k, v in orc.store.vars.items() if orc.store.trainable.get(k, False)}
            for v in params.values(): v.requires_grad_(True)
            loss, _ = orc.loss(x[i:i + 1], y[i:i + 1], iso=[isos[i]], cam=[cams[i]])
            loss.backward()
            gi = {k: (v.grad.numpy() / 6 if v.grad is not None else np.zeros(tuple(v.shape))) for k, v in params.items()}
            acc = gi if acc is None else {k: acc[k] + gi[k] for k in gi}
        for k in ("plain", "graph", "host0"):
            print("  %-13s worst rel diff vs oracle %.3e (%s)" % (k, *rel(res[k][1], acc)))
