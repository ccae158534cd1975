#!/usr/bin/env python
"""TEST INFRASTRUCTURE - writes tests/golden/<cfg>_network.json: the reaction networks of the BASELINE
configs compiled by vulcan_b200.network (our own JSON schema, not a copy of the reference text), so that
GPU-box tests (no /root/reference there) can rebuild the tables.  The species order / reactant lists in
these files are pinned against the reference's generated chem_funs (spec_list, re_wM_dict) by
tests/test_network.py through the <cfg>_static.npz fixtures."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from vulcan_b200.network import Network
REF = os.environ.get("VULCAN_REFERENCE", "/root/reference")
NETS = {"HD189": "NCHO_photo_network.txt", "Jupiter": "NCHO_photo_network_lowT_Jupiter.txt",
        "Earth": "NCHO_earth_photo_network.txt", "HD209S": "SNCHO_photo_network_2025.txt",
        "EarthS": "SNCHO_full_photo_network.txt",
        "HD189thermo": "NCHO_thermo_network.txt",
        "HD189cho": "CHO_photo_network.txt"}             # ni = 41: padded block size 48        # no photo section (use_photo = False)        # the network cfg_examples/vulcan_cfg_Earth.py names (ni = 99: padded block size 120)
# fixture-only ion test network: written into the scratch copy by oracle/stage_reference.py::write_ion_test_network
ION = "/tmp/vulcan_ref_HD189ion/thermo/NCHO_photo_ion_test_network.txt"
if os.path.exists(ION):
    NETS["HD189ion"] = ION
for tag, fn in NETS.items():
    net = Network.from_file(fn if os.path.isabs(fn) else os.path.join(REF, "thermo", fn))
    out = os.path.join(os.path.dirname(HERE), "tests", "golden", tag + "_network.json")
    with open(out, "w") as f:
        f.write(net.to_json())
    print(tag, net.ni, net.nr, out)
