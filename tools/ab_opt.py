"""In-process A/B of one library option on the 256-clip step (box-to-box variance exceeds most effects).
AB_OPT = option id (3 LN fuse, 5 last-block CLS, 1 attention impl), AB_MODES = comma-separated values."""
import runpy, os, sys
sys.argv = [os.path.join(os.path.dirname(os.path.abspath(__file__)), "ab_ln.py")]
runpy.run_path(sys.argv[0], run_name="__main__")
