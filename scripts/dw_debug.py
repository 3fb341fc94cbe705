"""Per-CTA timeline of the dW pass (NIW_DW_DEBUG=1) at the C2 size."""
import os, sys
os.environ["NIW_DW_DEBUG"] = "1"
sys.argv = [sys.argv[0]] + sys.argv[1:]
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_mlp.py")).read())
