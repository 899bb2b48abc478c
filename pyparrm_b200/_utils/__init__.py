"""Helpers the reference keeps under ``pyparrm._utils`` (only the device-accelerated ones)."""
