"""Device-side counterparts of the reference's ``codes/utils`` helpers that sit next to the hot path."""
