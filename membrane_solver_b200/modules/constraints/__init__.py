"""B200 twins of the reference constraint modules that provide constraint gradients on the hot path."""
