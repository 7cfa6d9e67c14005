"""Development (GPU): the Filon bench arm alone."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
print(json.dumps(bench.filon_arm(0), indent=1))
