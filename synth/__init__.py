"""Synthetic input generators of the shapes BASELINE.json names.

Bench / test infrastructure: deterministic, seeded, with the expected answer computed from the integer
columns BEFORE serialisation, so the text parsers are checked against an independent truth.
"""
