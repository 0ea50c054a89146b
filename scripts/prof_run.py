"""Phase profile of the persistent solver (needs a library built with MCT_PROFILE=1)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
from quick_bench import run
cases = [(1000, 148), (100, 8)]
if len(sys.argv) > 1:
    cases = [tuple(int(x) for x in a.split(":")) for a in sys.argv[1:]]
for Nk, team in cases:
    os.environ["MCT_TEAM_SIZE"] = str(team)
    run(Nk, 0.51)
