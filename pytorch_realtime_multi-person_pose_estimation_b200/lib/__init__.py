"""Host-side mirror of the reference's `lib` package for the inference hot path."""
