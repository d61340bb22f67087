"""Golden XDMF sidecar produced by the reference's own python-utils/writexmf.py (imported from /root/reference, which exists
only in the build container): tests/golden/xdmf_ref.xmf.  The inputs are rebuilt by tests/test_io.py."""
import os
import sys

sys.path.insert(0, "/root/reference/python-utils")
from writexmf import writexmf  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from xdmf_inputs import inputs  # noqa: E402


if __name__ == "__main__":
    x, y, z, ts, dt, names = inputs()
    writexmf(os.path.join(HERE, "xdmf_ref.xmf"), "double", x, y, z, ts, dt, names)
    print("wrote", os.path.join(HERE, "xdmf_ref.xmf"))
