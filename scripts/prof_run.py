import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__)))
from quick_bench import run
for Nk, team in ((1000, 148), (100, 8)):
    os.environ["MCT_TEAM_SIZE"] = str(team)
    run(Nk, 0.51)
