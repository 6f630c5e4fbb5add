"""Number-of-microbatches calculators (parity: megatron/microbatches.py: constant :42-58, ramp-up :61-144)."""
from __future__ import annotations

from abc import ABC, abstractmethod


def build_num_microbatches_calculator(args):
    if args.rampup_batch_size is None:
        calc = ConstantNumMicroBatches(args.global_batch_size, args.micro_batch_size, args.data_parallel_size)
        if args.rank == 0:
            print("setting number of micro-batches to constant {}".format(calc.get()), flush=True)
        return calc
    assert len(args.rampup_batch_size) == 3, \
        "expected the following format: --rampup_batch_size <start batch size> <batch size incerement> <ramp-up samples>"
    start, incr, samples = (int(v) for v in args.rampup_batch_size)
    if args.rank == 0:
        print("will use batch size rampup starting from global batch size {} to global batch size {} with batch "
              "size increments {} over {} samples.".format(start, args.global_batch_size, incr, samples), flush=True)
    return RampupBatchsizeNumMicroBatches(start, incr, samples, args.global_batch_size, args.micro_batch_size,
                                          args.data_parallel_size)


class NumMicroBatchesCalculator(ABC):
    def __init__(self):
        self.num_micro_batches = None
        self.current_global_batch_size = None

    def get(self):
        return self.num_micro_batches

    def get_current_global_batch_size(self):
        return self.current_global_batch_size

    @abstractmethod
    def update(self, consumed_samples, consistency_check):
        ...


class ConstantNumMicroBatches(NumMicroBatchesCalculator):
    def __init__(self, global_batch_size, micro_batch_size, data_parallel_size):
        super().__init__()
        per_step = micro_batch_size * data_parallel_size
        assert global_batch_size % per_step == 0, \
            "global batch size ({}) is not divisible by micro batch size ({}) times data parallel size ({})".format(
                global_batch_size, micro_batch_size, data_parallel_size)
        self.num_micro_batches = global_batch_size // per_step
        assert self.num_micro_batches >= 1
        self.current_global_batch_size = global_batch_size

    def update(self, consumed_samples, consistency_check):
        pass


class RampupBatchsizeNumMicroBatches(NumMicroBatchesCalculator):
    """Linear batch-size ramp: start, start+incr, ... up to global_batch_size over ``ramup_samples`` samples."""

    def __init__(self, start_batch_size, batch_size_increment, ramup_samples, global_batch_size, micro_batch_size,
                 data_parallel_size):
        super().__init__()
        self.micro_batch_size, self.data_parallel_size = micro_batch_size, data_parallel_size
        self.micro_batch_times_data_parallel_size = micro_batch_size * data_parallel_size
        assert self.micro_batch_times_data_parallel_size > 0
        assert start_batch_size > 0
        self.start_batch_size = start_batch_size
        assert global_batch_size > 0
        self.global_batch_size = global_batch_size
        diff = global_batch_size - start_batch_size
        assert diff >= 0
        assert batch_size_increment > 0
        self.batch_size_increment = batch_size_increment
        assert diff % batch_size_increment == 0, \
            "expected global batch size interval ({}) to be divisible by global batch size increment ({})".format(
                diff, batch_size_increment)
        num_increments = diff // batch_size_increment
        self.ramup_samples = ramup_samples
        assert self.ramup_samples >= 0
        self.rampup_samples_per_increment = self.ramup_samples / num_increments if num_increments else 0
        self.update(0, False)

    def update(self, consumed_samples, consistency_check):
        if consumed_samples > self.ramup_samples or self.rampup_samples_per_increment == 0:
            self.current_global_batch_size = self.global_batch_size
        else:
            steps = int(consumed_samples / self.rampup_samples_per_increment)
            self.current_global_batch_size = self.start_batch_size + steps * self.batch_size_increment
            assert self.current_global_batch_size <= self.global_batch_size
        if consistency_check:
            assert self.current_global_batch_size % self.micro_batch_times_data_parallel_size == 0, \
                "current global batch size ({}) is not divisible by micro-batch-size ({}) times data parallel size " \
                "({})".format(self.current_global_batch_size, self.micro_batch_size, self.data_parallel_size)
        self.num_micro_batches = self.current_global_batch_size // self.micro_batch_times_data_parallel_size
