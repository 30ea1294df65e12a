"""Generates the committed fixtures under tests/golden/.  Run in the build container:

    python tests/golden/make_golden.py

* kat_metrics.json  — the four distinct confusion matrices and the accuracy / F1 / recall /
  precision values printed next to them in the reference's own results file
  /root/reference/infos_training_train_on_batch.txt (:22-31, :51-60, :79-88, :107-116).
  These are the only numbers the reference pins for this path.
* oracle_small.npz  — seeded inputs/outputs of the CPU oracle on a small shape (regression pin
  for the oracle itself; the reference cannot be run here: TensorFlow is absent).
The reference tree is read here only; nothing at test / bench time touches /root/reference.
"""
import json
import os
import re
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def parse_results(path):
    txt = open(path).read()
    blocks = []
    for m in re.finditer(r"Confusion\s+matrix\s*\n(.*?)Accuracy:\s*([0-9.]+)\s*\nF1score:\s*\[(.*?)\]\s*\n"
                         r"Recall:\s*\[(.*?)\]\s*\nPrecision:\s*\[(.*?)\]", txt, re.S):
        rows = re.findall(r"\[+\s*([0-9\s]+?)\]", m.group(1))
        cm = [[int(v) for v in r.split()] for r in rows]
        f = lambda s: [float(v) for v in s.split()]
        blocks.append(dict(cm=cm, accuracy=float(m.group(2)), f1=f(m.group(3)), recall=f(m.group(4)),
                           precision=f(m.group(5)), src_offset=m.start()))
    return blocks


def main():
    ref = "/root/reference/infos_training_train_on_batch.txt"
    blocks = parse_results(ref)
    with open(os.path.join(HERE, "kat_metrics.json"), "w") as f:
        json.dump(dict(source="infos_training_train_on_batch.txt", blocks=blocks), f, indent=1)
    print("kat_metrics:", len(blocks), "blocks")

    from oracle import resuneta_oracle as O
    torch.manual_seed(0)
    out = {}
    for variant in ("v2", "v1"):
        p = O.init_params((64, 64, 3), 5, True, variant, seed=7)
        x, y = O.synth_batch(2, 64, 3, 5, seed=11, block=8)
        xt = torch.from_numpy(x)
        o = O.forward(p, xt, True, 5, True, variant)
        for k in ("seg", "color"):
            out[f"{variant}_train_{k}"] = o[k].numpy()[:, ::2, ::2]
        o = O.forward(p, xt, False, 5, True, variant)
        for k in ("seg", "color"):
            out[f"{variant}_infer_{k}"] = o[k].numpy()[:, ::2, ::2]
        yt = {k: torch.from_numpy(v) for k, v in y.items()}
        for k in o:
            out[f"{variant}_tanimoto_{k}"] = O.tanimoto_dual_loss(yt[k], o[k]).numpy()
    np.savez_compressed(os.path.join(HERE, "oracle_small.npz"), **out)
    print("oracle_small:", len(out), "arrays")


if __name__ == "__main__":
    main()
