"""CPU oracle for the MicKey hot path — TEST INFRASTRUCTURE, never imported by the product path."""
