"""Phase profile of one Nk=1000 solve at eta (argv, default 0.5155); needs MCT_LIB=.../libmct_sm100a_prof.so."""
import os, sys
sys.path.insert(0, os.path.dirname(__file__))
from quick_bench import run
os.environ.setdefault("MCT_TEAM_SIZE", "148")
run(1000, float(sys.argv[1]) if len(sys.argv) > 1 else 0.5155)
