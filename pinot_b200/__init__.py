"""pinot_b200 -- B200-native scan -> filter -> group-by-aggregate path for Apache Pinot segments."""
