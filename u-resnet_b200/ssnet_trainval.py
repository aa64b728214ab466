"""``ssnet_trainval``: train / inference driver, mirror of /root/reference/lib/ssnet_trainval.py.

Same public methods (``override_config``, ``initialize``, ``train_step``, ``ana``, ``ana_step``,
``batch_process``, ``reset``, ``num_class``, ``iterations``, ``current_iteration``,
``report_memory``, ``iteration_from_file_name``) and the same stdout report format.  What is
replaced because it cannot exist here:
  * larcv IO            -> ``synthetic_io.synthetic_threadio`` (same call surface)
  * tf.Session / graph  -> the uresnet_b200 CUDA library (``sess`` is ``None``)
  * tf.train.Saver      -> ``.npz`` files keyed by the TF variable names
                            (``<SAVE_FILE>-<iteration>.npz``; ``LOAD_FILE`` / ``AVOID_LOAD_PARAMS``
                            / iteration-from-suffix behave like ssnet_trainval.py:139-151)
  * tf.summary          -> ``LOGDIR/{train,test}``: TensorBoard event files with the reference's tags
                            (scalars ``accuracy_all`` / ``accuracy_nonzero`` / ``loss``, image
                            ``data_example`` for 2D, ssnet.py:81-89) when ``tensorboard`` is
                            importable, plus a ``scalars.jsonl`` copy of the scalars
  * larcv output        -> ``.npz`` per analysed batch next to ``ANA_OUTPUT_CONFIG``
Under ``torch.distributed`` (one process per GPU) every rank draws its own ``MINIBATCH_SIZE``
batch per minibatch (global batch = world x MINIBATCH_SIZE x NUM_MINIBATCHES) and the gradient
accumulators are all-reduced (SUM, ``dp.GradientExchange``) before Adam: the reference's
``NUM_MINIBATCHES`` accumulation spread over ranks (ssnet.py:76-79).
"""
from __future__ import annotations

import datetime
import glob
import json
import os
import sys
import time

import numpy as np

from .config import ssnet_config
from .synthetic_io import synthetic_threadio
from .uresnet import uresnet


class _summary_log(object):
    """tf.summary.FileWriter stand-in (ssnet_trainval.py:120-127, ssnet.py:81-89)."""

    def __init__(self, logdir):
        os.makedirs(logdir, exist_ok=True)
        self._jsonl = open(os.path.join(logdir, 'scalars.jsonl'), 'a')
        try:
            from torch.utils.tensorboard import SummaryWriter
            self._tb = SummaryWriter(logdir)
        except Exception:                 # tensorboard not installed: JSON lines only
            self._tb = None

    def add_summary(self, scalars, step, images=None):
        rec = dict(scalars)
        rec['step'] = int(step)
        self._jsonl.write(json.dumps(rec) + '\n')
        self._jsonl.flush()
        if self._tb is not None:
            for tag, value in scalars.items():
                self._tb.add_scalar(tag, value, step)
            if images is not None:        # tf.summary.image('data_example', grayscale_to_rgb(data), 10)
                img = np.asarray(images[:10], dtype=np.float32)
                lo, hi = float(img.min()), float(img.max())
                img = (img - lo) / (hi - lo) if hi > lo else np.zeros_like(img)
                self._tb.add_images('data_example', np.repeat(img, 3, axis=-1), step, dataformats='NHWC')
            self._tb.flush()

    def close(self):
        self._jsonl.close()
        if self._tb is not None:
            self._tb.close()


class ssnet_trainval(object):

    def __init__(self):
        self._cfg = ssnet_config()
        self._input_main = None
        self._input_test = None
        self._output = None
        self._iteration = -1
        self._sess = None
        self._rank, self._world = 0, 1
        self._saver_t0 = time.time()

    def __del__(self):
        try:
            self.reset()
        except Exception:
            pass

    def _report(self, metrics, descr):
        msg = ''
        for i, desc in enumerate(descr):
            if not desc:
                continue
            msg += '%s=%6.6f   ' % (desc, metrics[i])
        msg += '\n'
        sys.stdout.write(msg)
        sys.stdout.flush()

    def num_class(self):
        return self._cfg.NUM_CLASS

    def iteration_from_file_name(self, file_name):
        name = file_name[:-4] if file_name.endswith('.npz') else file_name
        return int((name.split('-'))[-1])

    def override_config(self, file_name):
        self._cfg.override(file_name)
        self._cfg.dump()

    def report_memory(self):
        import torch
        return float(torch.cuda.max_memory_allocated())

    # ------------------------------------------------------------------ IO
    def _make_io(self, name, cfg_file, batch, keywords, seed_offset=0):
        dims = list(self._cfg.SYNTHETIC_DIMS)
        if not dims:
            is3d = False
            if cfg_file and os.path.isfile(cfg_file):
                is3d = 'Tensor3D' in open(cfg_file).read()
            dims = [192, 192, 192, 1] if is3d else [512, 512, 1]
        io = synthetic_threadio()
        self._io_list = getattr(self, '_io_list', []) + [io]
        io.configure({'filler_name': name, 'verbosity': 0, 'filler_cfg': cfg_file, 'dims': dims,
                      'seed': self._cfg.SYNTHETIC_SEED + seed_offset, 'rank': self._rank,
                      'keyword_data': keywords[0], 'keyword_label': keywords[1],
                      'keyword_weight': keywords[2] or 'weight'})
        io.start_manager(batch)
        return io

    def initialize(self):
        if not self._cfg.MAIN_INPUT_CONFIG:
            print('Must provide larcv data filler configuration file!')
            return
        import torch
        import torch.distributed as dist
        if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not dist.is_initialized():
            torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
            dist.init_process_group('nccl')
        if dist.is_initialized():
            self._rank, self._world = dist.get_rank(), dist.get_world_size()

        cfg = self._cfg
        self._input_main = self._make_io('MainIO', cfg.MAIN_INPUT_CONFIG, cfg.MINIBATCH_SIZE,
                                         (cfg.KEYWORD_DATA, cfg.KEYWORD_LABEL, cfg.KEYWORD_WEIGHT))
        if cfg.TEST_INPUT_CONFIG:
            self._input_test = self._make_io('TestIO', cfg.TEST_INPUT_CONFIG, cfg.TEST_BATCH_SIZE,
                                             (cfg.KEYWORD_TEST_DATA, cfg.KEYWORD_TEST_LABEL, cfg.KEYWORD_TEST_WEIGHT),
                                             seed_offset=500000)
        if cfg.ANA_OUTPUT_CONFIG:
            self._output = os.path.splitext(cfg.ANA_OUTPUT_CONFIG)[0] + '_out'
            os.makedirs(self._output, exist_ok=True)

        self._input_main.next(store_entries=(not cfg.TRAIN), store_event_ids=(not cfg.TRAIN))
        dim_data = self._input_main.fetch_data(cfg.KEYWORD_DATA).dim()
        self._net = uresnet(dims=dim_data[1:], num_class=cfg.NUM_CLASS,
                            base_num_outputs=cfg.BASE_NUM_FILTERS, debug=cfg.DEBUG)
        if cfg.TRAIN:
            self._net.construct(trainable=cfg.TRAIN, use_weight=cfg.USE_WEIGHTS,
                                learning_rate=cfg.LEARNING_RATE, dtype=cfg.DTYPE, impl=cfg.KERNEL_IMPL,
                                seed=cfg.TF_RANDOM_SEED)
        else:
            self._net.construct(trainable=cfg.TRAIN, use_weight=cfg.USE_WEIGHTS, dtype=cfg.DTYPE,
                                impl=cfg.KERNEL_IMPL, seed=cfg.TF_RANDOM_SEED)
        for io in getattr(self, '_io_list', []):          # later batches land in page-locked buffers
            io._alloc = self._net.pinned_empty
        if self._world > 1 and cfg.TRAIN:
            from . import dp
            dp.broadcast_variables(self._net)
            ex = dp.attach(self._net)
            ex.expected_minibatches = cfg.NUM_MINIBATCHES

        self._writer_train = None
        self._writer_test = None
        if cfg.LOGDIR and self._rank == 0:
            self._writer_train = _summary_log(os.path.join(cfg.LOGDIR, 'train'))
            if self._input_test:
                self._writer_test = _summary_log(os.path.join(cfg.LOGDIR, 'test'))
        if cfg.SAVE_FILE:
            save_dir = cfg.SAVE_FILE[0:cfg.SAVE_FILE.rfind('/')] if '/' in cfg.SAVE_FILE else ''
            if save_dir and not os.path.isdir(save_dir):
                os.makedirs(save_dir, exist_ok=True)
            self._saver = True
        if cfg.LOAD_FILE:
            self._iteration = self.iteration_from_file_name(cfg.LOAD_FILE)
            path = cfg.LOAD_FILE if cfg.LOAD_FILE.endswith('.npz') else cfg.LOAD_FILE + '.npz'
            stored = np.load(path)
            values = {}
            # the reference walks tf.GraphKeys.MODEL_VARIABLES (ssnet_trainval.py:139-151): weights, beta and the
            # two moving statistics of every layer.  The statistics are part of the checkpoint (_save) but of no
            # computation -- nothing updates or reads them -- so "loading" them only reports
            names = []
            for name, _ in self._net.trainable_variables():
                names.append(name)
                if name.endswith('/BatchNorm/beta'):
                    names += [name[:-4] + 'moving_mean', name[:-4] + 'moving_variance']
            for name in names:
                if (name + ':0') in cfg.AVOID_LOAD_PARAMS or name in cfg.AVOID_LOAD_PARAMS:
                    print('\033[91mSkipping\033[00m loading variable', name + ':0', 'from input weight...')
                    continue
                print('\033[95mLoading\033[00m variable', name + ':0', 'from', cfg.LOAD_FILE)
                if not name.split('/')[-1].startswith('moving_'):
                    values[name] = stored[name]
            self._net.set_variables(values)
        self._batch_metrics = None
        self._descr_metrics = None

    # ---------------------------------------------------------------- train
    def _save(self):
        path = '%s-%d.npz' % (self._cfg.SAVE_FILE, self._iteration)
        v = self._net.get_variables()
        for name in list(v.keys()):               # TF checkpoints also carry the (never updated,
            if name.endswith('/BatchNorm/beta'):   # never read) moving statistics -- SURVEY 3.4
                v[name[:-4] + 'moving_mean'] = np.zeros_like(v[name])
                v[name[:-4] + 'moving_variance'] = np.ones_like(v[name])
        np.savez(path, **v)
        # tf.train.Saver(max_to_keep=CHECKPOINT_NMAX, keep_checkpoint_every_n_hours=CHECKPOINT_NHOUR)
        # (ssnet_trainval.py:133-135): the newest NMAX files are kept; a file that drops out of that
        # window is deleted unless NHOUR hours of training passed since the last one kept for good
        if not hasattr(self, '_kept_forever'):
            self._kept_forever, self._last_kept = set(), self._saver_t0      # Saver starts its clock at construction
        old = sorted((f for f in glob.glob(self._cfg.SAVE_FILE + '-*.npz') if f not in self._kept_forever),
                     key=os.path.getmtime)
        for f in old[:-max(1, int(self._cfg.CHECKPOINT_NMAX))]:
            nhour = float(self._cfg.CHECKPOINT_NHOUR)
            if nhour > 0 and os.path.getmtime(f) >= self._last_kept + nhour * 3600.0:
                self._kept_forever.add(f)
                self._last_kept = os.path.getmtime(f)
            else:
                os.remove(f)
        return path

    def _train_batch(self):
        """Current batch of the main stream; weights normalised per event ONCE per batch
        (ssnet_trainval.py:173)."""
        cfg = self._cfg
        io = self._input_main
        data = io.fetch_data(cfg.KEYWORD_DATA).data()
        label = io.fetch_data(cfg.KEYWORD_LABEL).data()
        weight = None
        if cfg.USE_WEIGHTS:
            weight = io.fetch_data(cfg.KEYWORD_WEIGHT or 'weight').data()
            if getattr(self, '_normalised_step', None) != (id(io), io._step):
                weight /= (np.sum(weight, axis=1).reshape([weight.shape[0], 1]))
                self._normalised_step = (id(io), io._step)
        return data, label, weight

    def train_step(self):
        cfg = self._cfg
        self._iteration += 1
        report_step = self._iteration % cfg.REPORT_STEPS == 0
        summary_step = cfg.SUMMARY_STEPS and (self._iteration % cfg.SUMMARY_STEPS) == 0
        checkpt_step = cfg.CHECKPOINT_STEPS and ((self._iteration + 1) % cfg.CHECKPOINT_STEPS) == 0

        self._net.zero_gradients(self._sess)
        for j in range(cfg.NUM_MINIBATCHES):
            minibatch_data, minibatch_label, minibatch_weight = self._train_batch()
            # the reference's sess.run blocks until the step is done and larcv's filler threads prepare
            # the next batch meanwhile (ssnet_trainval.py:176,187).  Same pipeline here, explicitly: the
            # kernels of minibatch j are only ENQUEUED, then the host produces minibatch j+1 and starts
            # its host->device transfer under them, and only then are the metrics of j fetched.
            self._net.accum_gradients(sess=self._sess, input_data=minibatch_data, input_label=minibatch_label,
                                      input_weight=minibatch_weight, fetch=False)
            self._input_main.next(store_entries=(not cfg.TRAIN), store_event_ids=(not cfg.TRAIN))
            self._net.stage_next(*self._train_batch())
            res, doc = self._net.fetch_metrics(minibatch_data.shape[0])
            if self._batch_metrics is None:
                self._batch_metrics = np.zeros((cfg.NUM_MINIBATCHES, len(res) - 1), dtype=np.float32)
                self._descr_metrics = doc[1:]
            self._batch_metrics[j, :] = res[1:]

        self._net.apply_gradients(self._sess)

        (test_data, test_label, test_weight) = (None, None, None)
        if (report_step or summary_step) and self._input_test:
            self._input_test.next()
            test_data = self._input_test.fetch_data(cfg.KEYWORD_TEST_DATA).data()
            test_label = self._input_test.fetch_data(cfg.KEYWORD_TEST_LABEL).data()
            if cfg.USE_WEIGHTS:
                test_weight = self._input_test.fetch_data(cfg.KEYWORD_TEST_WEIGHT or 'weight').data()
                test_weight /= (np.sum(test_weight, axis=1).reshape([test_weight.shape[0], 1]))

        if report_step and self._rank == 0:
            tstamp = datetime.datetime.fromtimestamp(time.time()).strftime('%Y-%m-%d %H:%M:%S')
            sys.stdout.write('@ iteration {:d} LR {:g} Mem {:g} @ {:s}\n'.format(
                self._iteration, self._net._opt._lr, self.report_memory(), tstamp))
            sys.stdout.write('Train set: ')
            self._report(np.mean(self._batch_metrics, axis=0), self._descr_metrics)
        if report_step and self._input_test:
            res, doc = self._net.run_test(self._sess, test_data, test_label, test_weight)
            if self._rank == 0:
                sys.stdout.write('Test set: ')
                self._report(res, doc)

        if summary_step:
            is2d = len(self._net._dims) == 3 and self._net._dims[-1] == 1      # ssnet.py:81-82

            def example(data):
                return np.asarray(data[:10]).reshape([-1] + list(self._net._dims)) if is2d else None
            s = self._net.make_summary(self._sess, minibatch_data, minibatch_label, minibatch_weight)
            if self._writer_train:
                self._writer_train.add_summary(s, self._iteration, example(minibatch_data))
            if self._input_test:
                s = self._net.make_summary(self._sess, test_data, test_label, test_weight)
                if self._writer_test:
                    self._writer_test.add_summary(s, self._iteration, example(test_data))

        if checkpt_step and self._cfg.SAVE_FILE and self._rank == 0:
            print('saved @', self._save())

    # ------------------------------------------------------------------ ana
    def ana(self, input_data, input_label=None):
        return self._net.inference(sess=self._sess, input_data=input_data, input_label=input_label)

    def ana_step(self, batch_mode=False):
        cfg = self._cfg
        self._iteration += 1
        batch_data = self._input_main.fetch_data(cfg.KEYWORD_DATA).data()
        batch_label = self._input_main.fetch_data(cfg.KEYWORD_LABEL).data()
        decision = None
        if batch_mode and self._output:
            # nothing but the shower/track map is kept: skip the softmax download altogether
            softmax = None
            decision, acc_all, acc_nonzero = self._net.decision(self._sess, batch_data, batch_label)
        elif self._output:
            softmax, acc_all, acc_nonzero, decision = self._net.inference(
                sess=self._sess, input_data=batch_data, input_label=batch_label, with_decision=True)
        else:
            softmax, acc_all, acc_nonzero = self.ana(input_data=batch_data, input_label=batch_label)

        copy_data, copy_label, copy_entries = (None, None, None)
        if not batch_mode:
            img_shape = list(softmax.shape)
            img_shape[-1] = -1
            copy_data = np.array(batch_data).reshape(img_shape)
            copy_label = np.array(batch_label).reshape(img_shape)
            copy_entries = np.array(self._input_main.fetch_entries())

        if self._output:
            entries = self._input_main.fetch_entries()
            # ssnet_trainval.py:269-287: (shower > track) + 2 (track >= shower), masked by data > 1,
            # evaluated on the device (urn_decision)
            for entry in range(len(decision)):
                print('Entry', entries[entry], 'Acc', acc_nonzero)
            np.savez_compressed(os.path.join(self._output, 'ssnet_%06d.npz' % self._iteration),
                                entries=np.asarray(entries), ssnet=decision)

        self._input_main.next(store_entries=(not cfg.TRAIN), store_event_ids=(not cfg.TRAIN))
        if not batch_mode:
            return {'entries': copy_entries, 'input': copy_data, 'label': copy_label,
                    'softmax': softmax, 'acc_all': acc_all, 'acc_nonzero': acc_nonzero}

    def batch_process(self):
        for i in range(self._cfg.ITERATIONS):
            if self._cfg.TRAIN and self._iteration >= self._cfg.ITERATIONS:
                print('Finished training (iteration %d)' % self._iteration)
                break
            if self._cfg.TRAIN:
                self.train_step()
            else:
                self.ana_step(batch_mode=True)

    def iterations(self):
        return self._cfg.ITERATIONS

    def current_iteration(self):
        return self._iteration

    def reset(self):
        if getattr(self, '_input_main', None) is not None:
            self._input_main.reset()
            self._input_main = None
        if getattr(self, '_input_test', None) is not None:
            self._input_test.reset()
            self._input_test = None
        for w in ('_writer_train', '_writer_test'):
            if getattr(self, w, None) is not None:
                getattr(self, w).close()
                setattr(self, w, None)
        self._output = None
