"""One C5 share (rank R of W, argv) with the phase timers on: prints {phase: ms} of the last call."""
import json, sys
sys.path.insert(0, '.')
from triumvirate_b200 import core
core.profile_enable(True)
sys.argv = [sys.argv[0]] + (sys.argv[1:3] if len(sys.argv) > 2 else ['0', '1']) + ['3']
exec(open('scripts/c5_share_once.py').read())
print("phases", json.dumps({k: round(v * 1e3, 2) for k, v in core.profile_report().items()}))
