"""metran_b200 -- B200-native batched Kalman filter / RTS smoother behind the Metran API."""
__version__ = "0.1.0"
