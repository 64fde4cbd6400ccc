"""Golden vectors for the hypothesis / trn / checkpoint glue, produced by running the UNMODIFIED
reference in the build container:
  * utils/text2trn.py and utils/average_checkpoints.py are executed as the scripts they are (subprocess);
  * src/utils/wer.py is imported (with a Levenshtein `editdistance` stand-in: the C extension is not installed);
  * the hypothesis line format is produced by the very statements of src/tasks/cassnat_task.py:346-355, extracted
    from the file's source text at run time (the decode loop is not importable as a function).
Writes tests/golden/glue.json (+ three tiny checkpoints as nested lists).  Run: python tests/golden/make_golden_glue.py"""
import json
import os
import subprocess
import sys
import tempfile
import textwrap
import types

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def lev(a, b):
    prev = list(range(len(b) + 1))
    for i, x in enumerate(a, 1):
        cur = [i]
        for j, y in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (x != y)))
        prev = cur
    return prev[-1]


def main():
    out = {}
    rng = np.random.default_rng(5)
    tmp = tempfile.mkdtemp()
    # ---- text2trn
    text = "1089-134686-0000 he hoped there would be stew\n1089-134686-0001\n2961-960-0003 a b\n"
    open(os.path.join(tmp, "text"), "w").write(text)
    subprocess.run([sys.executable, os.path.join(REF, "utils/text2trn.py"), os.path.join(tmp, "text"), os.path.join(tmp, "trn")], check=True)
    out["text"], out["trn"] = text, open(os.path.join(tmp, "trn")).read()
    # ---- average_checkpoints
    keys = ["encoder.layers.0.w", "module.att_generator.proj.bias", "scalar"]
    ckpts = []
    for ep in (3, 4, 5):
        st = {"encoder.layers.0.w": torch.from_numpy(rng.standard_normal((3, 4)).astype(np.float32)),
              "module.att_generator.proj.bias": torch.from_numpy(rng.standard_normal(5).astype(np.float32)),
              "scalar": torch.tensor(float(ep))}
        torch.save({"model_state": st, "epoch": ep}, os.path.join(tmp, f"model.{ep}.mdl"))
        ckpts.append({k: v.tolist() for k, v in st.items()})
    subprocess.run([sys.executable, os.path.join(REF, "utils/average_checkpoints.py"), "--exp_dir", tmp, "--out_name", "avg.mdl",
                    "--num", "3", "--last_epoch", "5"], check=True)
    avg = torch.load(os.path.join(tmp, "avg.mdl"), map_location="cpu")["model_state"]
    out["ckpts"], out["avg"] = ckpts, {k: avg[k].tolist() for k in keys}
    # ---- wer helpers
    ed = types.ModuleType("editdistance")
    ed.eval = lev
    sys.modules["editdistance"] = ed
    sys.path.insert(0, os.path.join(REF, "src"))
    import utils.wer as W
    B, T, U, V = 3, 12, 6, 9
    ctc_out = torch.from_numpy(rng.standard_normal((B, T, V)).astype(np.float32))
    att_out = torch.from_numpy(rng.standard_normal((B, U, V)).astype(np.float32))
    label = rng.integers(0, V, (B, U))
    label[0, 4:] = 0
    label[1, 5] = 2
    feat = [12, 9, 7]
    out["wer"] = {"ctc_out": ctc_out.tolist(), "att_out": att_out.tolist(), "label": label.tolist(), "feat_size": feat,
                  "ctc": [int(v) for v in W.ctc_greedy_wer(ctc_out, label, feat)],
                  "att": [int(v) for v in W.att_greedy_wer(att_out, label)]}
    # ---- hypothesis lines: the statements of cassnat_task.py:346-355, run on a fake batch
    src = open(os.path.join(REF, "src/tasks/cassnat_task.py")).read().split("\n")
    start = next(i for i, l in enumerate(src) if "#obtian hyp utt for seq" in l)
    end = next(i for i in range(start, len(src)) if "if args.print_utt2diff" in src[i])
    block = textwrap.dedent("\n".join(src[start + 1:end]))
    words = ["<blank>", "<sos>", "<eos>"] + [f"w{i}" for i in range(3, 12)]
    vocab = types.SimpleNamespace(word2index={"sos": 1, "eos": 2, "blank": 0}, index2word={i: w for i, w in enumerate(words)})
    utt_list = ["utt-a", "utt-b", "utt-c"]
    recog_results = [[{"hyp": [1, 5, 0, 7, 2, 9]}], [{"hyp": [1, 2]}], [{"hyp": [1, 3, 4, 11]}]]
    lines = []
    import io
    buf = io.StringIO()
    env = {"self": types.SimpleNamespace(vocab=vocab), "args": types.SimpleNamespace(padding_idx=0), "utt_list": utt_list,
           "recog_results": recog_results, "out_file": buf}
    exec(block, env)
    out["hyp"] = {"words": words, "utt_list": utt_list, "recog_results": recog_results, "lines": buf.getvalue()}
    json.dump(out, open(os.path.join(HERE, "glue.json"), "w"), indent=0)
    print("wrote glue.json:", {k: (len(v) if hasattr(v, "__len__") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
