"""Drop-in mirror of the reference's `src` package (same module paths and public names), so the reference's
`main.py` imports resolve against the sm_100a kernels.  See INTEGRATION.md."""
