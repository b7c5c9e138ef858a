[("helpers qry/qual/penalty","bt_core.cuh",186,218),("block load + lf","bt_core.cuh",221,257),("ftab/joined","bt_core.cuh",258,295),("bt_phase","bt_core.cuh",296,362),
 ("half_counts/partials","bt_core.cuh",363,403),("prologue","bt_core.cuh",404,422),("position","bt_core.cuh",424,493),
 ("rare: PHASE/BT_BEGIN","bt_core.cuh",496,544),("rare: FRAME_ENTER/POS","bt_core.cuh",545,574),("rare: BTLOOP","bt_core.cuh",575,645),("rare: FRAME_RET","bt_core.cuh",646,660),
 ("rare: CHILD_RET","bt_core.cuh",661,706),("rare: POS_END","bt_core.cuh",707,714),("rare: REPORT*","bt_core.cuh",715,797),("rare: BT_END","bt_core.cuh",798,811),
 ("begin/finish read","bt_core.cuh",812,831),("fast_iter","bt_core.cuh",832,904),("rare_iter loop","bt_core.cuh",905,920),("kernel loop","bt_lib.cu",86,178)]
