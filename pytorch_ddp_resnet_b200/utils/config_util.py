"""YAML run configuration (reference: resnet/utils/config_util.py:6-28). Unlike the reference's
ConfigParser, this one is a real dict, so `**config` works in-process as well as after pickling
(the reference only works through mp.spawn's pickle round trip)."""
from typing import Any, Dict, Optional

import yaml


class ConfigParser(dict):
    def __init__(self, defaults: Optional[Dict[str, Any]] = None) -> None:
        super().__init__()
        self.update(defaults or {})

    def read(self, config_path: str, verbose: bool = False) -> None:
        with open(config_path, "rb") as f:
            self.update(yaml.safe_load(f))
        if verbose:
            for k, v in self.items():
                print(f"{k}: {v}")

    def get(self, item: str) -> Any:  # KeyError on a missing key, like the reference
        return self[item]
