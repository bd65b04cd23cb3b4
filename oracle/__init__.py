"""ORACLE — test infrastructure only (see oracle/reference_math.py). Never imported by fer_vit_b200/."""
