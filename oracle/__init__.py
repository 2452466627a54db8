"""ORACLE -- test infrastructure only (see htdemucs_oracle.py header).

Nothing under ``demucs_b200/`` may import this package; tests/test_layout.py enforces it.
"""
