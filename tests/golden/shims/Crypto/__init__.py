"""Stand-in for the pycryptodome package (not installable here: no network), used ONLY by
tests/golden/make_model_traces.py to execute the reference's own tb/gcm_model.py.  It maps
the six calls that file makes (AES.new(..., MODE_GCM, nonce=), update, encrypt, decrypt,
digest, verify -- tb/gcm_model.py:18,22,26,30,35,44) onto `cryptography` (OpenSSL)."""
