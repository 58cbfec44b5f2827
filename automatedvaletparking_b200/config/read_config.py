"""read_config(config_name) -> dict, the reference's config loader API (config/read_config.py:17-23)."""
import os

import yaml


def read_config(config_name) -> dict:
    path = os.path.join(os.path.dirname(os.path.realpath(__file__)), config_name + '.yaml')
    with open(path, 'r', encoding='utf-8') as f:
        return yaml.load(f.read(), Loader=yaml.FullLoader)
