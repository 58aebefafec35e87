import sys, os
_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, _ROOT); sys.path.insert(0, os.path.join(_ROOT, "tests"))
import parity_cases as P
for (k, n, pro) in ((32, 152, 1), (114, 32, 2), (32, 32, 3), (86, 32, 0), (32, 86, 1), (64, 304, 1), (228, 64, 2), (128, 612, 1), (459, 128, 2)):
    for hw in ((128, 160), (64, 80), (256, 320)):
        for passes in (3,):
            try:
                r = P.case_pw_mma("cuda", k, n, hw=hw, prologue=pro, passes=passes, tol=(1, 1))
                print("K=%d N=%d pro=%d hw=%s: rel_l2 %.2e max %.2e" % (k, n, pro, hw, r[0], r[1]), flush=True)
            except Exception as e:
                print("FAIL", k, n, pro, hw, str(e)[:150], flush=True)
