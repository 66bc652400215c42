"""Dense-array mesh containers with the attribute surface the hot-path plugins touch."""
