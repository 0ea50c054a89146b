"""Config 2 (Nk=100, eta=0.51) against team size and fabric (development aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(__file__))
from quick_bench import run
for team, cl in ((1, 1), (2, 1), (4, 1), (8, 1), (11, 1), (16, 1), (11, 0), (24, 0), (37, 0)):
    os.environ["MCT_TEAM_SIZE"] = str(team); os.environ["MCT_CLUSTER"] = str(cl)
    print("cluster" if cl else "L2", end=" ")
    run(100, 0.51)
