"""Diagnostic (not a pytest): per-tap and per-shape conv3d error report, printed as a table."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "fast-cwdm_b200"), os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from test_conv3d_gpu import CASES, run_conv  # noqa: E402


def main():
    for case in CASES:
        try:
            got, ref = run_conv(*case)
            print("case", case, "max_err %.4g" % float((got - ref).abs().max()), "ref_max %.4g" % float(ref.abs().max()),
                  flush=True)
        except Exception:
            print("case", case, "EXC", traceback.format_exc(), flush=True)
            return
    for tap in range(27):
        def only(w, tap=tap):
            m = torch.zeros_like(w)
            m.view(w.shape[0], w.shape[1], 27)[:, :, tap] = 1
            return w * m
        got, ref = run_conv(1, 4, 16, 8, 64, 64, 3, use_bias=False, seed=tap, wfill=only)
        e = (got - ref).abs()
        print("tap", tap, (tap // 9, (tap // 3) % 3, tap % 3), "max_err %.4g" % float(e.max()),
              "frac_bad %.4f" % float((e > 0.05).float().mean()), flush=True)


if __name__ == "__main__":
    main()
