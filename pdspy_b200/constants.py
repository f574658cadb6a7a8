"""Constants the hot path shares with the reference (pdspy/constants/astronomy.py)."""
arcsec = 4.84813681e-6      # radians; the reference's truncated literal (astronomy.py:9), kept for parity
