"""Minimal Fortran namelist reader for the `&radsurf` / `&radsurf_driver` groups.

The reference reads both groups from one file with `read(unit, nml=...)`
(radsurf/radsurf_config.F90:125-247, driver/spartacus_surface_config.F90:76-165).
Only the syntax used by the shipped test namelists is supported: scalar
`key = value` entries separated by commas/newlines, `!` comments, logicals
written as true/false/.true./.false./T/F.
"""
import re


def _convert(text):
    t = text.strip().strip(",").strip()
    low = t.lower().strip(".")
    if low in ("true", "t"):
        return True
    if low in ("false", "f"):
        return False
    if (t.startswith("'") and t.endswith("'")) or (t.startswith('"') and t.endswith('"')):
        return t[1:-1]
    try:
        return int(t)
    except ValueError:
        pass
    return float(t.lower().replace("d", "e"))


def read_namelist(path):
    """Return {group_name: {key: value}} with lower-cased names."""
    groups = {}
    current = None
    with open(path) as fh:
        for raw in fh:
            line = raw.split("!", 1)[0].strip()
            if not line:
                continue
            if line.startswith("&"):
                current = line[1:].split()[0].lower()
                groups.setdefault(current, {})
                line = line[1 + len(current):].strip()
                if not line:
                    continue
            if line.startswith("/") or line.lower().startswith("&end"):
                current = None
                continue
            if current is None:
                continue
            for item in re.split(r",(?![^()]*\))", line):
                item = item.strip()
                if not item or item == "/":
                    continue
                if "=" not in item:
                    continue
                key, val = item.split("=", 1)
                if val.strip().endswith("/"):
                    val = val.strip()[:-1]
                groups[current][key.strip().lower()] = _convert(val)
    return groups
