"""Test-only CPU oracle (see crnn_oracle.py header). Never imported by the product package."""
