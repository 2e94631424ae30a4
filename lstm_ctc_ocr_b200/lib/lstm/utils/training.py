"""Sequence-equality accuracy (reference lib/lstm/utils/training.py:26-37)."""
from ..config import cfg


def accuracy_calculation(original_seq, decoded_seq, ignore_value=0, isPrint=True):
    if len(original_seq) != len(decoded_seq):
        print("original lengths is different from the decoded_seq,please check again")
        return 0
    hits = 0
    for i, truth in enumerate(original_seq):
        got = [int(j) for j in decoded_seq[i] if j != ignore_value]
        want = [int(l) for l in truth if l != ignore_value]
        if isPrint and i < cfg.VAL.PRINT_NUM:
            print("seq{0:4d}: origin: {1} decoded:{2}".format(i, list(truth), got))
        hits += int(want == got)
    return hits * 1.0 / len(original_seq)
