[("helpers qry/qual/penalty","bt_core.cuh",186,218),("block load + lf","bt_core.cuh",221,257),("ftab/joined","bt_core.cuh",258,295),("bt_phase","bt_core.cuh",296,362),
 ("half_counts/partials","bt_core.cuh",363,403),("prologue","bt_core.cuh",404,425),("position","bt_core.cuh",426,495),
 ("rare: PHASE/BT_BEGIN","bt_core.cuh",497,550),("rare: FRAME_ENTER/POS","bt_core.cuh",551,588),("rare: BTLOOP","bt_core.cuh",589,663),("rare: FRAME_RET","bt_core.cuh",664,682),
 ("rare: CHILD_RET","bt_core.cuh",683,732),("rare: POS_END","bt_core.cuh",733,744),("rare: REPORT*","bt_core.cuh",745,843),("rare: BT_END","bt_core.cuh",844,859),
 ("begin/finish read","bt_core.cuh",880,900),("fast_iter","bt_core.cuh",901,1002),("rare_iter sweep","bt_core.cuh",1003,1060),("kernel loop","bt_lib.cu",86,200)]
