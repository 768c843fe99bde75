"""ControllabilityMetrics and InceptionMetrics (reference metrics/metrics.py): the host logic of the reference - which face
model parameter is overwritten in which latent slice, which attribute columns are compared, how the scores are combined and
which files are written - around the B200 networks (generator, encoders, InceptionV3, MobileNetV2).  TensorBoard / AzureML
logging and the matplotlib plot are out of scope; the JSON / text files a resumed run or a
plotting script reads and the per-image PNG dumps of get_metrics are written as the reference writes them."""
import json
import os
import numpy as np

from .inception_distance import InceptionFeatureExtractor, compute_FID, compute_KID
from .celeba_attribute_prediction import CelebaAttributeClassifier
from .controllability_metric_configs import ControllabilityMetricConfigs
from .blendshape_names import blendshape_names


class ControllabilityMetrics:
    """Does setting a face-model parameter move the CelebA attribute it should, and nothing else?  For every attribute
    configuration two image sets are rendered from the same latents - the driven parameter's latent slice replaced by the
    synthetic encoder's embedding of the "set" value and of the "other" value - and the attribute classifier scores both."""

    def __init__(self, confinet_model, attribute_classifier, per_image_tuning_iters=0):
        self.confinet_model = confinet_model
        is_classifier = isinstance(attribute_classifier, CelebaAttributeClassifier) or hasattr(attribute_classifier, "predict_attributes")
        # anything else is the path of a saved classifier (metrics.py:22-24)
        self.attribute_classifier = attribute_classifier if is_classifier else CelebaAttributeClassifier.load(attribute_classifier)
        self.per_image_tuning_iters = per_image_tuning_iters
        if confinet_model is not None:
            self.facemodel_param_names = list(confinet_model.config["facemodel_inputs"].keys())

    # ---- face-model side
    def _latent_slice(self, param_name):
        """columns of the latent vector that belong to one face-model input (the inputs are laid out in config order)"""
        dims = [v[1] for v in self.confinet_model.config["facemodel_inputs"].values()]
        at = self.facemodel_param_names.index(param_name)
        start = int(np.sum(dims[:at]))
        return start, start + dims[at]

    def get_facemodel_params_for_config(self, attribute_config, other_param):
        """metrics.py:30-50: ONE sampled parameter set (a draw from the fitted distributions) whose driven input is
        overwritten - a vector as it is, a {blendshape name: value} dict on an otherwise zero blendshape vector"""
        params = self.confinet_model.sample_facemodel_params(1)
        value = attribute_config.facemodel_param_value_other if other_param else attribute_config.facemodel_param_value
        target = params[self.facemodel_param_names.index(attribute_config.facemodel_param_name)]
        if not isinstance(value, dict):
            target[:] = value
            return params
        if attribute_config.facemodel_param_name != "blendshape_values":
            raise NotImplementedError
        target[:] = 0
        for shape_name, amount in value.items():
            target[:, blendshape_names.index(shape_name)] = amount
        return params

    def get_images_for_controllable_attribute(self, attribute_config, latent_vectors, rotations, other_param=False):
        """metrics.py:52-67: every latent vector gets the driven input's slice of the encoded parameter set"""
        encoded = np.asarray(self.confinet_model.synthetic_encoder.predict(
            self.get_facemodel_params_for_config(attribute_config, other_param)))
        lo, hi = self._latent_slice(attribute_config.facemodel_param_name)
        edited = np.copy(latent_vectors)
        edited[:, lo:hi] = encoded[0, lo:hi]
        return self.confinet_model.generate_images(edited, rotations)

    def _render_all(self, latent_vectors, rotations):
        """-> (reconstruction, {config: images with the attribute}, {config: images with the other value}); the call order
        (reconstruction, then per configuration set / other) is the reference's - the face-model draws follow it"""
        model = self.confinet_model
        decoded = model.generate_images(latent_vectors, rotations)
        with_attr, without_attr = {}, {}
        for name, cfg in ControllabilityMetricConfigs.all_configs():
            with_attr[name] = self.get_images_for_controllable_attribute(cfg, latent_vectors, rotations)
            without_attr[name] = self.get_images_for_controllable_attribute(cfg, latent_vectors, rotations, other_param=True)
        return decoded, with_attr, without_attr

    def generate_images_for_metric(self, input_images):
        """metrics.py:69-102: latents from the encoder for the whole set at once, or (per_image_tuning_iters > 0) from
        fine_tune_on_img image by image, each image rendered with the generator fine-tuned on it"""
        if self.per_image_tuning_iters <= 0:
            latent_vectors, rotations = self.confinet_model.encode_images(input_images)
            return self._render_all(latent_vectors, rotations)
        names = [name for name, _ in ControllabilityMetricConfigs.all_configs()]
        decoded, with_attr, without_attr = [], {n: [] for n in names}, {n: [] for n in names}
        for img in input_images:
            latent_vectors, rotations = self.confinet_model.fine_tune_on_img(img[np.newaxis], n_iters=self.per_image_tuning_iters)
            one, one_with, one_without = self._render_all(latent_vectors, rotations)
            decoded.append(one[0])
            for n in names:
                with_attr[n].append(one_with[n][0])
                without_attr[n].append(one_without[n][0])
        return (np.array(decoded), {n: np.array(v) for n, v in with_attr.items()}, {n: np.array(v) for n, v in without_attr.items()})

    # ---- classifier side
    def get_metrics_for_attribute_pairs(self, set_attributes, not_set_attributes, attribute_config):
        """metrics.py:104-130 -> (mean of the driven attribute when set, when set to the other value, mean absolute
        difference of the attributes that should stay constant, correlation of the driven attribute with the setting)"""
        names = self.attribute_classifier.config["predicted_attributes"]
        driven = names.index(attribute_config.driven_attribute)
        may_move = set(attribute_config.ignored_attributes) | {attribute_config.driven_attribute}
        fixed = [i for i, n in enumerate(names) if n not in may_move]
        count = len(set_attributes)
        assert count == len(not_set_attributes)
        on, off = set_attributes[:, driven], not_set_attributes[:, driven]
        correlation = np.corrcoef(np.vstack((np.hstack((np.ones(count), np.zeros(count))), np.hstack((on, off)))))[0, 1]
        drift = np.mean(np.mean(np.abs(set_attributes[:, fixed] - not_set_attributes[:, fixed]), axis=0))
        return float(np.mean(on)), float(np.mean(off)), float(drift), float(correlation)

    def get_metrics_for_attribute_config(self, attribute_config, images_with_attribute, images_without_attribute):
        predict = self.attribute_classifier.predict_attributes
        return self.get_metrics_for_attribute_pairs(predict(images_with_attribute), predict(images_without_attribute), attribute_config)

    def get_metrics_from_attribute_images(self, images_with_attributes, images_without_attributes):
        """metrics.py:158-169: the four numbers per configuration, their means, and the scalar the training logs plot"""
        metrics = {name: self.get_metrics_for_attribute_config(cfg, images_with_attributes[name], images_without_attributes[name])
                   for name, cfg in ControllabilityMetricConfigs.all_configs()}
        means = tuple(np.mean(list(metrics.values()), axis=0))
        metrics["contr_attribute_means"] = means
        metrics["controllability"] = 10 * means[2] + (1 - means[0])          # weights "based on perceived importance"
        return metrics

    def get_metrics(self, input_images, img_output_dir=None):
        """metrics.py:139-156; with ``img_output_dir`` every input, its reconstruction and the two images of every attribute
        configuration are written as PNG files under the reference's names (OpenCV, imported on demand)"""
        decoded, with_attr, without_attr = self.generate_images_for_metric(input_images)
        if img_output_dir is not None:
            import cv2
            os.makedirs(img_output_dir, exist_ok=True)
            put = lambda stem, i, img: cv2.imwrite(os.path.join(img_output_dir, "%s_%04d.png" % (stem, i)), np.asarray(img))
            for i in range(len(input_images)):
                put("gt_img", i, input_images[i])
                put("raw_img", i, decoded[i])
                for name in with_attr:
                    put(name + "_img", i, with_attr[name][i])
                    put(name + "_img_not_set", i, without_attr[name][i])
        return self.get_metrics_from_attribute_images(with_attr, without_attr)

    def update_and_log_metrics(self, images, metrics_dict, output_dir, aml_run=None, tb_log_writer=None):
        """metrics.py:171-199: appends to metrics_dict and rewrites controllability_metrics.json"""
        os.makedirs(output_dir, exist_ok=True)
        fresh = self.get_metrics(images)
        for key, value in fresh.items():
            metrics_dict.setdefault(key, []).append(value)
            if aml_run is not None:
                aml_run.log(key, value)
        with open(os.path.join(output_dir, "controllability_metrics.json"), "w") as fp:
            json.dump({key: metrics_dict[key] for key in fresh}, fp, indent=4)


class InceptionMetrics:
    def __init__(self, confignet_config, dataset, n_samples_for_metrics=1000, weights=None, device=None):
        """metrics.py:201-207: one np.random.randint draw of n_samples_for_metrics rows of the dataset's precomputed
        Inception features (dataset.inception_features), before anything else touches the NumPy stream"""
        self.n_samples_for_metrics = n_samples_for_metrics
        self.inception_feature_extractor = InceptionFeatureExtractor(confignet_config["output_shape"], weights=weights, device=device)
        metric_sample_idxs = np.random.randint(0, dataset.imgs.shape[0], n_samples_for_metrics)
        self.gt_inception_features = np.asarray(dataset.inception_features)[metric_sample_idxs]

    def get_metrics(self, generated_images):
        generated = self.inception_feature_extractor.get_features(generated_images)
        kid = compute_KID(generated, self.gt_inception_features)
        fid = compute_FID(generated, self.gt_inception_features)
        return kid, fid

    def update_and_log_metrics(self, images, metrics_dict, output_dir, aml_run=None, tb_log_writer=None):
        """metrics.py:215-264: appends kid / fid and rewrites inception_metrics.txt (step, kid, fid per row)"""
        os.makedirs(output_dir, exist_ok=True)
        kid, fid = self.get_metrics(images)
        metrics_dict.setdefault("kid", []).append(kid)
        metrics_dict.setdefault("fid", []).append(fid)
        assert len(metrics_dict["kid"]) == len(metrics_dict["fid"])
        if "training_step_number" in metrics_dict:
            steps = metrics_dict["training_step_number"]
            assert len(steps) == len(metrics_dict["kid"])
        else:
            steps = range(len(metrics_dict["kid"]))
        if aml_run is not None:
            aml_run.log("Kernel Inception Distance", kid)
            aml_run.log("Frechet Inception Distance", fid)
        header = "\t".join(["step_number", "kid", "fid"])
        table = np.stack((steps, metrics_dict["kid"], metrics_dict["fid"]), axis=1)
        np.savetxt(os.path.join(output_dir, "inception_metrics.txt"), table, header=header)
