"""CPU oracle (test infrastructure).  See anchor_pipeline_oracle.py."""
