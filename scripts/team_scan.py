"""Solve time of config 3 against the team size (development aid): fewer CTAs with full row blocks can need fewer unit slots."""
import os, sys
sys.path.insert(0, os.path.dirname(__file__))
from quick_bench import run
for team in [int(a) for a in sys.argv[1:]] or [148, 144, 140, 136, 132, 128, 126, 125, 124, 120, 112]:
    os.environ["MCT_TEAM_SIZE"] = str(team)
    run(1000, 0.5155)
