"""Generates tests/golden/pr_kitti08.npz by running the REFERENCE's own scripts/pr_mpe.py:get_points_ours2 (imported from
/root/reference, never copied) on the reference's own sample files:
    sample_data/ts-sens_pose-kitti08.txt  +  results/outcome_txt/outcome-kitti08.txt
matplotlib is not installed here and is only used by the script's plotting half, so an empty stub module stands in for it.
The fixture stores the parsed inputs (pose translations, outcome columns) and the reference's outputs (PR points, max-F1,
similarity threshold, TP pose errors) so that tests/test_eval.py can pin contour_context_b200/eval.py:pr_metrics on a box
where /root/reference does not exist.
    python tests/golden/make_golden_pr.py
"""
import contextlib
import io
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))


def main():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, os.path.join(REF, "scripts"))
    import pr_mpe  # the reference script

    fp_pose = os.path.join(REF, "sample_data", "ts-sens_pose-kitti08.txt")
    fp_out = os.path.join(REF, "results", "outcome_txt", "outcome-kitti08.txt")
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        plots = pr_mpe.get_points_ours2(fp_pose, fp_out)
    log = buf.getvalue()

    def grab(pat):
        return float(re.search(pat, log).group(1))

    from contour_context_b200 import eval as ev

    tf, it, isr, corr, err = ev.read_outcome(fp_out)
    poses = pr_mpe.get_gt_sens_poses(fp_pose)
    m = re.search(r"Max F1 score: ([0-9.eE+-]+) @(\d+)", log)
    np.savez_compressed(
        os.path.join(HERE, "pr_kitti08.npz"),
        gt_xyz=poses[:, [3, 7, 11]].astype(np.float64), tfpn=tf.astype(np.int8), id_tgt=it.astype(np.int32), id_src=isr.astype(np.int32),
        corr=corr, err=err,
        ref_pr_sorted=plots[0],  # (recall, precision) sorted by recall, as the script plots them
        ref_max_f1=float(m.group(1)), ref_f1_idx=int(m.group(2)), ref_sim_thres=grab(r"sim thres for Max F1 score: ([0-9.eE+-]+)"),
        ref_tp_count=int(grab(r"TP count:\s+([0-9]+)")), ref_rot_mean_deg=grab(r"Rot mean err:\s+([0-9.eE+-]+)"),
        ref_rot_rmse_deg=grab(r"Rot rmse\s+:\s+([0-9.eE+-]+)"), ref_trans_mean=grab(r"Trans mean err:\s+([0-9.eE+-]+)"),
        ref_trans_rmse=grab(r"Trans rmse\s+:\s+([0-9.eE+-]+)"), ref_log=np.array(log))
    print(log[-600:])


if __name__ == "__main__":
    main()
