"""OptiSpeech: the public module (reference: optispeech/model/optispeech.py:12-154).

Same constructor (`dim, generator, vocoder, discriminator, train_args, data_args, inference_args, optimizer,
scheduler` — the sub-models arrive as partials, as Hydra's `_partial_: true` produces them), same methods:
`synthesise(InferenceInputs) -> InferenceOutputs`, `prepare_input(text, ...)`, `training_step`,
`configure_optimizers`, `load_from_checkpoint`.
"""
from __future__ import annotations

from typing import Any, Dict

import torch

from ..values import InferenceInputs, InferenceOutputs
from .base_module import BaseModule


class OptiSpeech(BaseModule):
    def __init__(self, dim, generator, vocoder, discriminator, train_args, data_args, inference_args, optimizer=None,
                 scheduler=None):
        super().__init__()
        self.save_hyperparameters(dict(dim=dim, generator=generator, vocoder=vocoder, discriminator=discriminator,
                                       train_args=train_args, data_args=data_args, inference_args=inference_args,
                                       optimizer=optimizer, scheduler=scheduler))
        if (train_args.gradient_accumulate_batches is not None) and (train_args.gradient_accumulate_batches <= 0):
            raise ValueError("gradient_accumulate_batches should be a positive number")
        if data_args.num_speakers < 1:
            raise ValueError("num_speakers should be a positive integer >= 1")

        self.train_args = train_args
        self.data_args = data_args
        self.inference_args = inference_args
        self.text_processor = self.data_args.text_processor
        self.num_speakers = data_args.num_speakers
        self.sample_rate = data_args.feature_extractor.sample_rate
        self.hop_length = data_args.feature_extractor.hop_length
        self.automatic_optimization = False

        num_languages = getattr(self.text_processor, "num_languages", 1) if self.text_processor is not None else 1
        self.generator = generator(
            dim=dim,
            vocoder=vocoder,
            feature_extractor=data_args.feature_extractor,
            data_statistics=data_args.data_statistics,
            num_speakers=self.data_args.num_speakers,
            num_languages=num_languages,
        )
        self.discriminator = discriminator(feature_extractor=data_args.feature_extractor)

    @torch.inference_mode()
    def synthesise(self, inputs: InferenceInputs) -> InferenceOutputs:
        inputs = inputs.as_torch().to(self.device)
        out = self.generator.synthesise(x=inputs.x, x_lengths=inputs.x_lengths, sids=inputs.sids, lids=inputs.lids,
                                        d_factor=inputs.d_factor, p_factor=inputs.p_factor, e_factor=inputs.e_factor)
        return InferenceOutputs(wav=out["wav"], wav_lengths=out["wav_lengths"], durations=out["durations"], pitch=out["pitch"],
                                energy=out["energy"], latency=out["latency"], rtf=out["rtf"], am_rtf=out["am_rtf"], v_rtf=out["v_rtf"])

    # the README of the reference spells it `synthesize` (README.md:90); keep both
    synthesize = synthesise

    def prepare_input(self, text: str, *, language: str | None = None, speaker: str | int | None = None, d_factor: float = None,
                      p_factor: float = None, e_factor: float = None, split_sentences: bool = True) -> InferenceInputs:
        """Reference :83-154.  Text normalisation / phonemisation is the text processor's job (CPU, out of scope here);
        any object with the reference TextProcessor interface (`__call__(text, lang, split_sentences)`, `languages`,
        `is_multi_language`) works."""
        if self.text_processor is None:
            raise RuntimeError("this model was built without a text_processor; pass phoneme ids with InferenceInputs.from_ids_and_lengths")
        languages = self.text_processor.languages
        if language is None:
            language = languages[0]
        sid = None
        if self.num_speakers > 1:
            if speaker is None:
                sid = 0
            elif isinstance(speaker, str):
                try:
                    sid = self.speakers.index(speaker)
                except (ValueError, IndexError, AttributeError):
                    raise ValueError(f"A speaker with the given name `{speaker}` was not found in speaker list")
            elif isinstance(speaker, int):
                sid = speaker
        lid = None
        if self.text_processor.is_multi_language:
            try:
                lid = languages.index(language)
            except (ValueError, IndexError):
                raise ValueError(f"A language with the given name `{language}` was not found in language list")
        input_ids, clean_text = self.text_processor(text, lang=language, split_sentences=split_sentences)
        if split_sentences:
            lengths = [len(ids) for ids in input_ids]
        else:
            lengths, input_ids = [len(input_ids)], [input_ids]
        sids = [sid] * len(input_ids) if sid is not None else None
        lids = [lid] * len(input_ids) if lid is not None else None
        ia = self.inference_args
        inputs = InferenceInputs.from_ids_and_lengths(
            ids=input_ids, lengths=lengths, clean_text=clean_text, sids=sids, lids=lids,
            d_factor=d_factor or ia.d_factor, p_factor=p_factor or ia.p_factor, e_factor=e_factor or ia.e_factor)
        return inputs.as_torch().to(self.device)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict: bool = True, resume_training: bool = False,
                             **overrides) -> "OptiSpeech":
        """Lightning-style checkpoint: {'state_dict', 'hyper_parameters', 'epoch', 'global_step', 'optimizer_states',
        'lr_schedulers', ...} (reference optispeech/infer.py:38, optispeech/train.py:84-91).  Checkpoints written by the
        reference itself pickle OmegaConf containers and live text-processor / feature-extractor objects in
        `hyper_parameters`; they are read through `optispeech_b200.checkpoint.load_checkpoint_file`, which resolves the
        `optispeech.*` class paths to this implementation and OmegaConf containers to plain dict / list / namespace objects
        (OmegaConf is not a dependency).  `resume_training=True` also restores optimizer, scheduler and step counters — what
        `Trainer.fit(ckpt_path=...)` does for the reference."""
        from ..checkpoint import hyper_parameters_from_checkpoint, load_checkpoint_file

        ckpt: Dict[str, Any] = load_checkpoint_file(checkpoint_path, map_location=map_location or "cpu")
        hparams = hyper_parameters_from_checkpoint(ckpt)
        hparams.update(overrides)
        model = cls(**hparams)
        model.load_state_dict(ckpt["state_dict"], strict=strict)
        model.on_load_checkpoint(ckpt)
        if map_location is not None and str(map_location) != "cpu":
            model = model.to(map_location)
        if resume_training:
            model.load_training_state(ckpt)
        return model

    def load_training_state(self, ckpt: Dict[str, Any]) -> None:
        """Optimizer moments / step counts (torch.optim.AdamW layout, so the reference's own `optimizer_states` load as well),
        LR schedulers and the batch counter behind `global_step` (base_lightning_module.py:294-303)."""
        opts, scheds = self.optimizers(), self.lr_schedulers()
        for opt, sd in zip(opts, ckpt.get("optimizer_states") or []):
            opt.load_state_dict(sd)
        for sched, sd in zip(scheds, ckpt.get("lr_schedulers") or []):
            sched.load_state_dict(sd)
            for group, lr in zip(sched.optimizer.param_groups, sched.get_last_lr()):
                group["lr"] = lr
        acc = self.train_args.gradient_accumulate_batches or 1
        self._fit.total_batch_idx = int(ckpt.get("global_step", 0)) * int(acc)

    def save_checkpoint(self, path, epoch: int = 0, global_step: int | None = None):
        """Writes the dictionary layout `load_from_checkpoint` (and Lightning) reads, optimizer and scheduler state included."""
        ckpt = {"state_dict": self.state_dict(), "hyper_parameters": vars(self.hparams), "epoch": epoch,
                "global_step": self.global_step if global_step is None else global_step}
        if self._optimizers is not None:
            ckpt["optimizer_states"] = [o.state_dict() for o in self._optimizers]
            ckpt["lr_schedulers"] = [s.state_dict() for s in self._schedulers]
        torch.save(ckpt, path)
