"""Drop-in `diffusion` package (same import names as /root/reference/diffusion)."""
