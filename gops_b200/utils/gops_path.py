"""Name helpers of the plugin registries (reference: gops/utils/gops_path.py:12-30)."""
import os

gops_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
algorithm_path = os.path.join(gops_path, "algorithm")
apprfunc_path = os.path.join(gops_path, "apprfunc")
env_path = os.path.join(gops_path, "env")


def underline2camel(s: str, first_upper: bool = False) -> str:
    """fhadp -> FHADP (first_upper), pyth_lq_model -> LqModel style used by the registries."""
    parts = s.split("_")
    head = parts.pop(0).upper() if first_upper else ""
    return head + "".join(p[0].upper() + p[1:] for p in parts)


def camel2underline(s: str) -> str:
    out = ""
    for i in range(len(s) - 1):
        out += ("_" + s[i].lower()) if (s[i].isupper() and s[i + 1].islower()) else s[i].lower()
    return out + s[-1].lower()
