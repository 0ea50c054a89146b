"""Single-equation latency at Nk=500 vs team size (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(__file__))
from quick_bench import run
for team, cl in ((16, 1), (16, 0), (24, 0), (37, 0), (50, 0), (74, 0), (100, 0), (148, 0)):
    os.environ["MCT_TEAM_SIZE"] = str(team); os.environ["MCT_CLUSTER"] = str(cl)
    print("cluster" if cl else "L2", end=" ")
    run(500, 0.51)
