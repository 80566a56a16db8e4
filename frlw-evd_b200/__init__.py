"""B200-native event-representation encoders."""
