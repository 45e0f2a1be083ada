"""TEST INFRASTRUCTURE — writes tests/golden/maps.json: what the reference's OWN parsers (utils.parse_office_world,
utils.parse_map_emoji, /root/reference/multiagent_rlrm/utils/utils.py:169-364) return for the reference's OWN layouts
(config_office.py, config_frozen_lake.py). tests/test_host_logic.py compares this repo's ASCII maps and parsers with it.
Needs /root/reference (run in the build container):   python oracle/gen_maps_fixture.py
"""
import json
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(_HERE, "ref_shim"), os.environ.get("RLRM_REFERENCE_ROOT", "/root/reference")]

from multiagent_rlrm.environments.frozen_lake.config_frozen_lake import config as fc  # noqa: E402
from multiagent_rlrm.environments.office_world.config_office import config as oc  # noqa: E402
from multiagent_rlrm.utils.utils import parse_map_emoji, parse_office_world  # noqa: E402


def main():
    out = {"office_world": {}, "frozen_lake": {},
           "generator": "oracle/gen_maps_fixture.py; reference parse_office_world / parse_map_emoji (utils.py:169-364)"}
    for name, m in oc["maps"].items():
        c, g, w = parse_office_world(m["layout"])
        out["office_world"][name] = {"coordinates": {k: [list(p) for p in v] for k, v in c.items()},
                                     "goals": {k: list(v) for k, v in g.items()}, "walls": [[list(a), list(b)] for a, b in w],
                                     "grid_size": list(m["grid_size"]), "start": list(m["agents"][0]["position"])}
    for name, m in fc["maps"].items():
        h, g, d = parse_map_emoji(m["layout"])
        out["frozen_lake"][name] = {"holes": [list(p) for p in h], "goals": {k: list(v) for k, v in g.items()}, "dims": list(d)}
    path = os.path.join(os.path.dirname(_HERE), "tests", "golden", "maps.json")
    json.dump(out, open(path, "w"))
    print("wrote", path)


if __name__ == "__main__":
    main()
