"""Known-answer vectors transcribed from the reference's own tests (file:line provenance on every block).

All paths are under /root/reference/pinot-core/src/test/java/org/apache/pinot/queries/.
"""

# BaseSingleValueQueriesTest.java:99-104
FILTER = (" WHERE column1 > 100000000"
          " AND column3 BETWEEN 20000000 AND 1000000000"
          " AND column5 = 'gFuH'"
          " AND (column6 < 500000000 OR column11 NOT IN ('t', 'P'))"
          " AND daysSinceEpoch = 126164076")

# InnerSegmentAggregationSingleValueQueriesTest.java:29-40
AGGREGATION_QUERY = "SELECT COUNT(*), SUM(column1), MAX(column3), MIN(column6), AVG(column7) FROM testTable"
SMALL_GROUP_BY = " GROUP BY column9"                                                      # ARRAY_BASED
MEDIUM_GROUP_BY = " GROUP BY column9, column11, column12"                                 # INT_MAP_BASED
LARGE_GROUP_BY = " GROUP BY column1, column6, column9, column11, column12"                # LONG_MAP_BASED
VERY_LARGE_GROUP_BY = (" GROUP BY column1, column3, column6, column7, column9, column11, column12, column17, "
                       "column18")                                                        # ARRAY_MAP_BASED

# (query, stats(docsScanned, entriesInFilter, entriesPostFilter, totalDocs),
#  expected (count, sum(column1), max(column3), min(column6), avg.sum(column7), avg.count))
INNER_SEGMENT_AGGREGATION = [
    # :43-51
    (AGGREGATION_QUERY, (30000, 0, 120000, 30000), (30000, 32317185437847, 2147419555, 1689277, 28175373944314, 30000)),
    # :53-59
    (AGGREGATION_QUERY + FILTER, (6129, 63064, 24516, 30000),
     (6129, 6875947596072, 999813884, 1980174, 4699510391301, 6129)),
]

# (query, regime, stats, group key VALUES, expected as above)
INNER_SEGMENT_GROUP_BY = [
    # testSmallAggregationGroupBy :95-110
    (AGGREGATION_QUERY + SMALL_GROUP_BY, "ARRAY", (30000, 0, 150000, 30000), (11270,),
     (1, 815409257, 1215316262, 1328642550, 788414092, 1)),
    (AGGREGATION_QUERY + FILTER + SMALL_GROUP_BY, "ARRAY", (6129, 63064, 30645, 30000), (242920,),
     (3, 4348938306, 407993712, 296467636, 5803888725, 3)),
    # testMediumAggregationGroupBy :113-131
    (AGGREGATION_QUERY + MEDIUM_GROUP_BY, "INT_MAP", (30000, 0, 210000, 30000), (1813102948, "P", "HEuxNvH"),
     (4, 2062187196, 1988589001, 394608493, 4782388964, 4)),
    (AGGREGATION_QUERY + FILTER + MEDIUM_GROUP_BY, "INT_MAP", (6129, 63064, 42903, 30000),
     (1176631727, "P", "KrNxpdycSiwoRohEiTIlLqDHnx"), (1, 716185211, 489993380, 371110078, 487714191, 1)),
    # testLargeAggregationGroupBy :134-153
    (AGGREGATION_QUERY + LARGE_GROUP_BY, "LONG_MAP", (30000, 0, 210000, 30000),
     (484569489, 16200443, 1159557463, "P", "MaztCmmxxgguBUxPti"), (2, 969138978, 995355481, 16200443, 2222394270, 2)),
    (AGGREGATION_QUERY + FILTER + LARGE_GROUP_BY, "LONG_MAP", (6129, 63064, 42903, 30000),
     (1318761745, 353175528, 1172307870, "P", "HEuxNvH"), (2, 2637523490, 557154208, 353175528, 2427862396, 2)),
    # testVeryLargeAggregationGroupBy :156-174
    (AGGREGATION_QUERY + VERY_LARGE_GROUP_BY, "ARRAY_MAP", (30000, 0, 270000, 30000),
     (1784773968, 204243323, 628170461, 1985159279, 296467636, "P", "HEuxNvH", 402773817, 2047180536),
     (1, 1784773968, 204243323, 628170461, 1985159279, 1)),
    (AGGREGATION_QUERY + FILTER + VERY_LARGE_GROUP_BY, "ARRAY_MAP", (6129, 63064, 55161, 30000),
     (1361199163, 178133991, 296467636, 788414092, 1719301234, "P", "MaztCmmxxgguBUxPti", 1284373442, 752388855),
     (1, 1361199163, 178133991, 296467636, 788414092, 1)),
]

# InterSegmentAggregationSingleValueQueriesTest.java -- 2 identical segments x 2 servers == 4 identical segments.
# GROUP_BY :39 : " GROUP BY column9 ORDER BY v1 DESC, v2 DESC LIMIT 1" (v1, v2 = the two aggregations)
INTER_GROUP_BY = " GROUP BY column9"
# (select list, ascending?, [no filter, filter, group-by top row, filter + group-by top row])  -- final values.
# testMin orders its group-by "ORDER BY v1, v2 LIMIT 1" (:138); all others use GROUP_BY (:39, DESC).
INTER_SEGMENT = [
    # testMax :92-118
    ("MAX(column1), MAX(column3)", False, [(2146952047.0, 2147419555.0), (2146952047.0, 999813884.0),
                                    (2146952047.0, 2146630496.0), (2146952047.0, 999813884.0)]),
    # testMin :121-148
    ("MIN(column1), MIN(column3)", True, [(240528.0, 17891.0), (101116473.0, 20396372.0), (240528.0, 17891.0),
                                    (101116473.0, 91804599.0)]),
    # testSum :151-175
    ("SUM(column1), SUM(column3)", False, [(129268741751388.0, 129156636756600.0), (27503790384288.0, 12429178874916.0),
                                    (69526727335224.0, 69225631719808.0), (19058003631876.0, 8606725456500.0)]),
    # testAvg :178-202
    ("AVG(column1), AVG(column3)", False, [(1077239514.5949, 1076305306.305), (1121871038.68037, 506982332.96280),
                                    (2142595699.0, 334963174.0), (2142595699.0, 334963174.0)]),
    # testDistinctCount :235-259
    ("DISTINCTCOUNT(column1), DISTINCTCOUNT(column3)", False, [(6582, 21910), (1872, 4556), (3495, 11961), (1272, 3289)]),
]
# testCount :47-70: COUNT(*) = 120000; with FILTER 24516; top group by COUNT(*) DESC: 64420 / 17080
INTER_SEGMENT_COUNT = [120000, 24516, 64420, 17080]
# merged ExecutionStatistics for the FILTER case: docsScanned 24516, entriesInFilter 252256 (= 4 x 63064)
INTER_SEGMENT_FILTER_STATS = (24516, 252256)

# InterSegmentGroupBySingleValueQueriesTest.java:43-110 (column11 values are strings)
INTER_GROUP_BY_COLUMN11_SUM = [("", 5935285005452.0), ("P", 88832999206836.0), ("gFuH", 63202785888.0),
                               ("o", 18105331533948.0), ("t", 16331923219264.0)]
