"""Kernel-parameter holder and Decider -- host-side mirror of GNNAdvisor/param.py.

`InputProperty` carries what the layer code reads (row_pointers, column_index, degrees, partPtr,
part2Node, partSize, dimWorker, warpPerBlock: gnn_conv.py:15-19,36-46,105-112) and reproduces the
reference Decider's arithmetic (param.py:51-120) so the same graph gets the same
(partSize, dimWorker, warpPerBlock) in auto mode.  `inputProperty` is an alias with the
reference's spelling.

What the three numbers mean to the B200 kernel (csrc/aggregate.cu):
    partSize      neighbours per group = one unit of work
    dimWorker     lanes that cooperate on one neighbour row (rounded down to a power of two and
                  capped by what the row needs); the other lanes of the warp take other groups
    warpPerBlock  warps per CTA (capped at 16)
"""
import math


class InputProperty(object):
    MAX_WARP_PER_BLOCK = 8      # param.py:44
    SMEM_FRACTION = 0.4         # param.py:45
    SMEM_GAP = 100              # param.py:46

    def __init__(self, row_pointers=None, column_index=None, degrees=None,
                 partSize=None, dimWorker=None, warpPerBlock=None, sharedMem=None,
                 hiddenDim=None, dataset_obj=None, enable_rabbit=False, manual_mode=True, verbose=False):
        if dataset_obj is None:
            raise ValueError("Dataset object MUST SET !!!")      # param.py:15-16
        self.dataset_obj = dataset_obj
        self.row_pointers, self.column_index, self.degrees = row_pointers, column_index, degrees
        self.num_nodes = dataset_obj.num_nodes
        self.avgNodeDegree = dataset_obj.avg_degree
        self.avgEdgeSpan = dataset_obj.avg_edgeSpan
        self.partSize, self.dimWorker, self.warpPerBlock = partSize, dimWorker, warpPerBlock
        self.dimWorker_input = self.dimWorker_hidden = dimWorker
        self.warpPerBlock_input = self.warpPerBlock_hidden = warpPerBlock
        self.inputDim = dataset_obj.num_features
        self.hiddenDim = hiddenDim
        self.manual_mode, self.enable_rabbit, self.verbose_flag = manual_mode, enable_rabbit, verbose
        self.state_set_input = False
        self.reorder_status = False
        self.MAX_warpPerBlock = self.MAX_WARP_PER_BLOCK
        self.share_memory = (sharedMem if sharedMem is not None else 100) * self.SMEM_FRACTION
        self.gap_smem = self.SMEM_GAP
        self.partPtr = None
        self.part2Node = None

    # ------------------------------------------------------------------ Decider (param.py:51-120)
    def decider(self):
        ds = self.dataset_obj
        if self.manual_mode:
            # manual: the CLI values are used for both layers; reorder iff asked (param.py:58-70)
            ds.reorder_flag = bool(self.enable_rabbit)
            if self.enable_rabbit:
                ds.rabbit_reorder()
                self.row_pointers, self.column_index = ds.row_pointers, ds.column_index
            self.reorder_status = bool(self.enable_rabbit)
            self._say("\n=> MANUAL Config Complete !!!\n")
            return self

        self.partSize = int(self.avgNodeDegree)                                    # :73
        per_group = self.partSize * 4
        est_in = self.MAX_warpPerBlock * (per_group + self.inputDim * 4 + self.gap_smem * 4) / 1e3      # :75
        # the hidden-layer estimate really uses hiddenDim, not hiddenDim*4, in the reference (:82)
        est_hid = self.MAX_warpPerBlock * (per_group + self.hiddenDim + 4 * self.gap_smem) / 1e3
        smem_in = min(est_in, self.share_memory)
        smem_hid = min(est_hid, self.share_memory)
        self._say("input-layer shared memory (KB): {:.3f} ".format(est_in))
        self._say("input-layer updated (KB): {:.3f}".format(smem_in))
        self._say("hidden-layer shared memory (KB): {:.3f}".format(est_hid))
        self._say("hidden-layer updated (KB): {:.3f}".format(smem_hid))
        self.warpPerBlock_input = min(int(smem_in * 1e3 / (per_group + self.inputDim * 4)), self.MAX_warpPerBlock)     # :89,92
        self.warpPerBlock_hidden = min(int(smem_hid * 1e3 / (per_group + self.hiddenDim * 4)), self.MAX_warpPerBlock)  # :90,93
        self.dimWorker_input = 32 if self.inputDim > 32 else self.inputDim          # :96-99
        self.dimWorker_hidden = 32 if self.hiddenDim > 32 else self.hiddenDim       # :102-105
        if self.enable_rabbit:
            # reorder iff sqrt(avg edge span) > sqrt(N)/100 (:110); like the reference, the CSR held by
            # THIS object is not refreshed in auto mode
            want = math.sqrt(self.avgEdgeSpan) > math.sqrt(self.num_nodes) / 100
            ds.reorder_flag = want
            self.reorder_status = want
            ds.rabbit_reorder()
        self._say("\n=> AUTO Decider Complete !!!\n")
        return self

    # ------------------------------------------------------------------ Decider re-tuned for B200
    @staticmethod
    def b200_choice(avg_degree, dim):
        """(partSize, dimWorker, warpPerBlock) for csrc/aggregate.cu on B200, from the parameter studies in
        profiles/r01_params_*.txt (the reference's s7-4_1 / s7-4_2 sweeps re-run on the new kernel):
          * partSize: 16..64 is flat and best; partSize = avgDegree (the reference's rule, param.py:73) costs
            9 % on Reddit (491) and 3x on sparse graphs (2..4).  32 for avgDegree >= 24, else 16.
          * dimWorker: lanes per neighbour row.  Half the row's 16-byte chunks (two chunks per lane, twice as
            many groups in flight per warp) is 7 % faster than one chunk per lane when HBM-bound and equal
            when L2-bound; never fewer than 4 lanes.
          * warpPerBlock: 2..4 is best (smaller CTAs retire and refill faster); 4."""
        part_size = 32 if avg_degree >= 24 else 16
        chunks = (int(dim) + 3) // 4
        need = 1
        while need < chunks and need < 32:
            need *= 2
        dim_worker = max(4, need // 2) if need >= 16 else max(min(need, 4), need)
        return part_size, dim_worker, 4

    def decider_b200(self):
        """Auto mode with the B200 choices instead of the reference's sm_86 heuristics; same fields are set."""
        self.partSize, self.dimWorker_input, self.warpPerBlock_input = self.b200_choice(self.avgNodeDegree, self.inputDim)
        _, self.dimWorker_hidden, self.warpPerBlock_hidden = self.b200_choice(self.avgNodeDegree, self.hiddenDim)
        self.dimWorker, self.warpPerBlock = self.dimWorker_hidden, self.warpPerBlock_hidden
        if self.enable_rabbit:
            want = math.sqrt(self.avgEdgeSpan) > math.sqrt(self.num_nodes) / 100      # the reference's test (:110)
            self.dataset_obj.reorder_flag = want
            self.reorder_status = want
            self.dataset_obj.rabbit_reorder()
            if want:   # unlike the reference the reordered CSR is picked up in auto mode too
                self.row_pointers, self.column_index = self.dataset_obj.row_pointers, self.dataset_obj.column_index
        self._say("\n=> B200 Decider Complete !!!\n")
        return self

    # ------------------------------------------------------------------ per-layer switch (param.py:122-141)
    def set_input(self):
        self.dimWorker, self.warpPerBlock = self.dimWorker_input, self.warpPerBlock_input
        self.state_set_input = True
        return self

    def set_hidden(self):
        self.dimWorker, self.warpPerBlock = self.dimWorker_hidden, self.warpPerBlock_hidden
        self.state_set_input = False
        return self

    def print_param(self):
        if not self.verbose_flag:
            return
        mode = "manual" if self.manual_mode else "auto"
        layer = "INPUT" if self.state_set_input else "HIDDEN"
        print("# {} {} partSize: {}".format(mode, layer, self.partSize))
        print("# {} {} dimWorker: {}".format(mode, layer, self.dimWorker))
        print("# {} {} warpPerBlock: {}".format(mode, layer, self.warpPerBlock))
        if not self.manual_mode:
            print("# {} {} reorder_flag: {}".format(mode, layer, self.reorder_status))

    def _say(self, msg):
        if self.verbose_flag:
            print(msg)


inputProperty = InputProperty
