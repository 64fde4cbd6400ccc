"""Hypothesis / scoring glue around the decode call and checkpoint handling: what the reference's decode
loop and helper scripts do with the `batch_top_seqs` the path returns and with trained checkpoints.  Host-side, no GPU.

Reference code mirrored (paths relative to /root/reference):
  hypothesis line format "<utt> tok tok ..."           src/tasks/cassnat_task.py:346-355 (+ print_utt2diff :357-365)
  text -> sclite .trn                                   utils/text2trn.py:10-18
  checkpoint averaging                                  utils/average_checkpoints.py:22-45
  test-time checkpoint loading ("module." tolerant)     src/tasks/base_task.py:44-54
  greedy CTC / attention token error counts             src/utils/wer.py:7-52
"""
import os

import torch

from .sharding import edit_distance


def hypothesis_tokens(hyp_ids, vocab, padding_idx=0):
    """Token strings of one hypothesis: sos and padding ids are skipped, eos ends it (cassnat_task.py:348-353)."""
    sos, eos = vocab.word2index["sos"], vocab.word2index["eos"]
    out = []
    for idx in hyp_ids:
        if idx == sos or idx == padding_idx:
            continue
        if idx == eos:
            break
        out.append(vocab.index2word[idx])
    return out


def hypothesis_lines(utt_list, recog_results, vocab, padding_idx=0):
    """One "<utt> tok tok ..." line per utterance from the best hypothesis of `beam_decode`'s result."""
    return [utt + " " + " ".join(hypothesis_tokens(res[0]["hyp"], vocab, padding_idx))
            for utt, res in zip(utt_list, recog_results)]


def length_diff_lines(utt_list, recog_results, labels):
    """The optional utt2diff file (cassnat_task.py:357-365): hypothesis length minus label length, clipped to +-3."""
    lines = []
    for utt, res, lab in zip(utt_list, recog_results, labels):
        diff = len(res[0]["hyp"]) - len(lab)
        diff = 3 if diff > 2 else (-3 if diff < -2 else diff)
        lines.append(utt + " " + str(diff))
    return lines


def write_hypotheses(out_file, utt_list, recog_results, vocab, padding_idx=0):
    """Appends the batch's hypothesis lines to an open text file (what the decode loop prints, flushed per line)."""
    for line in hypothesis_lines(utt_list, recog_results, vocab, padding_idx):
        print(line, flush=True, file=out_file)


def text_to_trn(text_path, trn_path):
    """Kaldi text ("<utt> words...") -> sclite trn ("words... (<utt with - as _>)"), utils/text2trn.py."""
    with open(text_path, "r") as rf, open(trn_path, "w") as wf:
        line = rf.readline()
        while line:
            parts = line.strip().split(" ", 1)
            if len(parts) > 1:
                wf.write(parts[1] + " (" + parts[0].replace("-", "_") + ")\n")
            else:
                wf.write(" (" + parts[0].replace("-", "_") + ")\n")
            line = rf.readline()


def average_checkpoints(exp_dir, last_epoch, num=12, out_name=None):
    """Average of the `model_state` of model.<epoch>.mdl for the last `num` epochs ending at `last_epoch`
    (utils/average_checkpoints.py:22-45).  Returns the averaged state dict; saves {'model_state': avg} when out_name
    is given.  Integer entries are averaged with the reference's in-place true division semantics (float result
    stored back where the dtype allows; integer tensors raise in torch, exactly as they would there)."""
    avg = None
    for epoch in range(last_epoch - num + 1, last_epoch + 1):
        states = torch.load(os.path.join(exp_dir, "model.{}.mdl".format(epoch)), map_location=torch.device("cpu"))["model_state"]
        if avg is None:
            avg = states
        else:
            for k in avg.keys():
                avg[k] += states[k]
    for k in avg.keys():
        if avg[k] is not None:
            avg[k] /= num
    if out_name is not None:
        torch.save({"model_state": avg}, os.path.join(exp_dir, out_name))
    return avg


def load_test_model(model, resume_model):
    """Copies a checkpoint's `model_state` into the model's parameters, accepting DDP's "module." key prefix
    (src/tasks/base_task.py:44-54).  Works with the reference's checkpoints: parameter names are identical."""
    checkpoint = torch.load(resume_model, map_location="cpu") if isinstance(resume_model, (str, os.PathLike)) else resume_model
    model_state = checkpoint["model_state"]
    with torch.no_grad():
        for name, param in model.named_parameters():
            if name not in model_state:
                name = "module." + name
            param.data.copy_(model_state[name])
    return model


def ctc_greedy_wer(ctc_out, label, feat_size, pad=0):
    """(errors, tokens) of the greedy CTC path against the labels (src/utils/wer.py:7-29)."""
    pred = torch.max(ctc_out, dim=-1)[1].cpu().numpy()
    errs = toks = 0
    for i in range(pred.shape[0]):
        p_seq = []
        for j in range(int(feat_size[i])):
            if pred[i][j] != 0 and not (j != 0 and pred[i][j] == pred[i][j - 1]):
                p_seq.append(int(pred[i][j]))
        h_seq = [int(t) for t in label[i] if t != pad and t != 2]
        errs += edit_distance(p_seq, h_seq)
        toks += len(h_seq)
    return errs, toks


def att_greedy_wer(att_out, label, pad=0):
    """(errors, tokens) of the greedy attention output against the labels (src/utils/wer.py:31-52)."""
    pred = torch.max(att_out, dim=-1)[1].cpu().numpy()
    errs = toks = 0
    for i in range(pred.shape[0]):
        p_seq = []
        for j in range(pred.shape[1]):
            if pred[i][j] == pad and label[i][j] == 1:
                continue
            if pred[i][j] == 2:
                break
            p_seq.append(int(pred[i][j]))
        h_seq = [int(t) for t in label[i] if t != pad and t != 2]
        errs += edit_distance(p_seq, h_seq)
        toks += len(h_seq)
    return errs, toks
