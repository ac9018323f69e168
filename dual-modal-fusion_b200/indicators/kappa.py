"""OA / AA / Kappa from a confusion matrix M[pred][target] — same names and return values as the
reference's indicators/kappa.py:10-22, 69-84.  O(C^2) float64 on the host; all sums are exact
integers below 2^53, so the results are bit-identical to the reference for an identical matrix."""
import json
import os

import numpy as np


def kappa(matrix):
    M = np.asarray(matrix, dtype=np.float64)
    n = np.sum(M)
    sum_po, sum_pe = 0, 0
    for i in range(len(M[0])):
        sum_po += M[i][i]
        sum_pe += np.sum(M[i, :]) * np.sum(M[:, i])
    po, pe = sum_po / n, sum_pe / (n * n)
    return (po - pe) / (1 - pe)


def aa_oa(matrix):
    """-> [aa, oa, kappa, rows]; class 0 (background) is left out of the per-class accuracies and of
    the OA numerator, but not of the OA denominator (reference: indicators/kappa.py:71-82)."""
    M = np.asarray(matrix, dtype=np.float64)
    per_target = np.sum(M, axis=0)
    accuracy, on_display, correct = [], [], 0
    with np.errstate(invalid='ignore', divide='ignore'):
        for i in range(1, M.shape[0]):
            a = M[i][i] / per_target[i]
            correct += M[i][i]
            accuracy.append(a)
            on_display.append([per_target[i], M[i][i], a])
            print("Category:{}. Overall:{}. Correct:{}. Accuracy:{:.6f}".format(i, per_target[i], M[i][i], a))
        aa = np.mean(accuracy)
        oa = correct / np.sum(per_target, axis=0)
        k = kappa(M)
    print("OA:{:.6f} AA:{:.6f} Kappa:{:.6f}".format(oa, aa, k))
    return [aa, oa, k, on_display]


def expo_result(result, cfg, time, group_num):
    """Report writer (reference: indicators/kappa.py:87-118 writes an xlsx through openpyxl).  The
    spreadsheet is reporting only; when openpyxl is missing the same numbers go to a JSON file next
    to where the xlsx would be."""
    aa, oa, k, rows = result
    record = {'group': group_num, 'AA': float(aa), 'OA': float(oa), 'Kappa': float(k),
              'train_time_s': float(time[0]), 'test_time_s': float(time[1]),
              'per_class': [[float(v) for v in r] for r in rows]}
    path = cfg['RESULT_excel']
    os.makedirs(os.path.dirname(path) or '.', exist_ok=True)
    try:
        from openpyxl import Workbook, load_workbook
    except ImportError:
        jpath = os.path.splitext(path)[0] + '.json'
        old = json.load(open(jpath)) if os.path.exists(jpath) else []
        json.dump(old + [record], open(jpath, 'w'), indent=1)
        return jpath
    wb = load_workbook(path) if os.path.exists(path) else Workbook()
    ws = wb.active
    col = group_num * 8 + 1
    ws.cell(1, col, 'Category'); ws.cell(1, col + 1, 'Overall'); ws.cell(1, col + 2, 'Correct'); ws.cell(1, col + 3, 'Accuracy')
    for i, r in enumerate(rows):
        ws.cell(i + 2, col, i + 1)
        for j, v in enumerate(r):
            ws.cell(i + 2, col + 1 + j, float(v))
    base = len(rows) + 3
    for j, (name, v) in enumerate([('OA', oa), ('AA', aa), ('KAPPA', k), ('Train time(s)', time[0]), ('Test time(s)', time[1])]):
        ws.cell(base + j, col, name)
        ws.cell(base + j, col + 1, float(v))
    wb.save(path)
    return path
