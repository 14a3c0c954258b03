"""Run a reference Examples script UNMODIFIED (runpy, cwd = its directory) against this repo's drop-in packages.

  python tools/run_unmodified_script.py SCRIPT [--seconds S] [--max-prints N] [--random JSON] [--stub-matplotlib]

The harness only controls the script's environment, never its text:
  * sys.path: repo root first (PDP / JinEnv / casadi packages of this repo); --stub-matplotlib adds tools/stubs;
  * --random '{"random": [[..]], "rand": [[..]], "randn": [[..]]}': the first calls of numpy.random.random / rand / randn return
    these arrays (so a stored trial's initial parameter can be reproduced), later calls fall through to NumPy;
  * the run ends (exit code 0) after S seconds or after the script has printed N lines.
Test infrastructure (tests/test_gpu_dropin_run.py, tests/test_dropin_scripts.py)."""
import argparse
import builtins
import json
import os
import runpy
import signal
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Done(BaseException):
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("script")
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--max-prints", type=int, default=0)
    ap.add_argument("--random", default="")
    ap.add_argument("--stub-matplotlib", action="store_true")
    args = ap.parse_args()
    script = os.path.abspath(args.script)
    sys.path.insert(0, ROOT)
    if args.stub_matplotlib:
        try:
            import matplotlib  # noqa: F401
        except ImportError:
            sys.path.insert(1, os.path.join(ROOT, "tools", "stubs"))
    import numpy as np
    if args.random:
        queues = {k: [np.asarray(v, dtype=np.float64) for v in vs] for k, vs in json.loads(args.random).items()}
        for name, q in queues.items():
            orig = getattr(np.random, name)

            def patched(*a, _q=q, _orig=orig, **k):
                if _q:
                    v = _q.pop(0)
                    want = _orig(*a, **k)
                    return v.reshape(np.shape(want)) if np.size(v) == np.size(want) else v
                return _orig(*a, **k)
            setattr(np.random, name, patched)
    count = [0]
    real_print = builtins.print

    def counting_print(*a, **k):
        real_print(*a, **k)
        sys.stdout.flush()
        count[0] += 1
        if args.max_prints and count[0] >= args.max_prints:
            raise _Done()
    builtins.print = counting_print

    def on_alarm(signum, frame):
        raise _Done()
    signal.signal(signal.SIGALRM, on_alarm)
    signal.setitimer(signal.ITIMER_REAL, args.seconds)
    os.chdir(os.path.dirname(script))
    sys.argv = [script]
    try:
        runpy.run_path(script, run_name="__main__")
    except _Done:
        pass
    finally:
        signal.setitimer(signal.ITIMER_REAL, 0)
        builtins.print = real_print
    sys.stdout.flush()
    os._exit(0)


if __name__ == "__main__":
    main()
