"""Stand-in for cocotb: tb/gcm_model.py only imports its logger (tb/gcm_model.py:2)."""
import logging

log = logging.getLogger("cocotb-shim")
