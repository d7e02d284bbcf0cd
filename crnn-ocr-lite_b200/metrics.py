"""Edit-distance evaluation of the reference (utils.py:262-298; predict.py:183-191)."""
import numpy as np


def levenshtein(seq1, seq2):
    """Classic O(n*m) edit distance with a rolling row (the reference fills the full numpy matrix)."""
    n, m = len(seq1), len(seq2)
    prev = list(range(m + 1))
    for i in range(1, n + 1):
        cur = [i] + [0] * m
        a = seq1[i - 1]
        for j in range(1, m + 1):
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (a != seq2[j - 1]))
        prev = cur
    return float(prev[m])


def edit_distance(y_pred, y_true):
    n = len(y_true)
    return sum(levenshtein(a, b) / n for a, b in zip(y_pred, y_true))


def normalized_edit_distance(y_pred, y_true):
    n = len(y_true)
    return sum(levenshtein(a, b) / (len(b) * n) for a, b in zip(y_pred, y_true))
