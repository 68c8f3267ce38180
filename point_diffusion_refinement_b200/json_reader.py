"""Experiment-config reader: the reference stores lists as strings and eval()s them back
(pointnet2/json_reader.py:5-24).  Here they are parsed with ``ast.literal_eval`` (no code execution);
the resulting dict is what the reference's models receive as ``pointnet_config``."""
import ast
import json


def restore_string_to_list_in_a_dict(d):
    for k, v in d.items():
        if isinstance(v, str) and v.lstrip().startswith("["):
            try:
                d[k] = ast.literal_eval(v)
            except (ValueError, SyntaxError):
                pass
        elif isinstance(v, dict):
            restore_string_to_list_in_a_dict(v)
    return d


def replace_list_with_string_in_a_dict(d):
    for k, v in d.items():
        if isinstance(v, list):
            d[k] = str(v)
        elif isinstance(v, dict):
            replace_list_with_string_in_a_dict(v)
    return d


def read_config(path):
    with open(path) as f:
        return restore_string_to_list_in_a_dict(json.load(f))
