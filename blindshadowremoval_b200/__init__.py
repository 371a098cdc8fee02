"""B200-native forward pass of the BlindShadowRemoval generator (GSC + TSM variants)."""
