import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in range(2):
    for vote in ("1", "0"):
        env = dict(os.environ, DXM_VOTE=vote, GRIDS="-4", KINDS="voce")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/sweep_grid.py")], capture_output=True, text=True, env=env)
        print("voce vote", vote, r.stdout.strip()[:120])
        env = dict(os.environ, DXM_VOTE=vote)
        r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/bench_fefp.py"), "4e7", "3e-2"], capture_output=True, text=True, env=env)
        print("fefp vote", vote, r.stdout.strip()[:130])
