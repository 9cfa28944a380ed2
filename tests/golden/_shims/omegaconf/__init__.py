"""Import shim (omegaconf is not installed); UNetModel.__init__ only does an isinstance-style
check against ListConfig (openai_unetmodel.py:476)."""
