"""CPU oracle package — test infrastructure, never imported by the product path."""
