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
    def __init__(self, confinet_model, attribute_classifier, per_image_tuning_iters=0):
        self.confinet_model = confinet_model
        if isinstance(attribute_classifier, CelebaAttributeClassifier) or hasattr(attribute_classifier, "predict_attributes"):
            self.attribute_classifier = attribute_classifier
        else:       # a path: metrics.py:22-24
            self.attribute_classifier = CelebaAttributeClassifier.load(attribute_classifier)
        self.per_image_tuning_iters = per_image_tuning_iters
        if confinet_model is not None:
            self.facemodel_param_names = list(self.confinet_model.config["facemodel_inputs"].keys())

    def get_facemodel_params_for_config(self, attribute_config, other_param):
        """metrics.py:30-50: one sampled face-model parameter set with the configured parameter overwritten"""
        facemodel_params = self.confinet_model.sample_facemodel_params(1)
        param_value = attribute_config.facemodel_param_value_other if other_param else attribute_config.facemodel_param_value
        param_idx = self.facemodel_param_names.index(attribute_config.facemodel_param_name)
        if isinstance(param_value, dict):
            if attribute_config.facemodel_param_name != "blendshape_values":
                raise NotImplementedError
            facemodel_params[param_idx][:] = 0
            for key, value in param_value.items():
                facemodel_params[param_idx][:, blendshape_names.index(key)] = value
        else:
            facemodel_params[param_idx][:] = param_value
        return facemodel_params

    def get_images_for_controllable_attribute(self, attribute_config, latent_vectors, rotations, other_param=False):
        """metrics.py:52-67: the latent slice of the driven parameter is replaced in every latent vector"""
        facemodel_params = self.get_facemodel_params_for_config(attribute_config, other_param)
        latent_vector_with_attribute_set = np.asarray(self.confinet_model.synthetic_encoder.predict(facemodel_params))
        modified_param_idx = self.facemodel_param_names.index(attribute_config.facemodel_param_name)
        facemodel_param_dims = list(self.confinet_model.config["facemodel_inputs"].values())
        start_idx = int(np.sum([x[1] for x in facemodel_param_dims[:modified_param_idx]]))
        end_idx = start_idx + facemodel_param_dims[modified_param_idx][1]
        modified_latent_vectors = np.copy(latent_vectors)
        modified_latent_vectors[:, start_idx:end_idx] = latent_vector_with_attribute_set[0, start_idx:end_idx]
        return self.confinet_model.generate_images(modified_latent_vectors, rotations)

    def generate_images_for_metric(self, input_images):
        """metrics.py:69-102"""
        configs = ControllabilityMetricConfigs.all_configs()
        if self.per_image_tuning_iters > 0:
            raw_decoded_images = []
            images_with_attributes = {name: [] for name, _ in configs}
            images_without_attributes = {name: [] for name, _ in configs}
            for img in input_images:
                img = img[np.newaxis]
                latent_vectors, rotations = self.confinet_model.fine_tune_on_img(img, n_iters=self.per_image_tuning_iters)
                raw_decoded_images.append(self.confinet_model.generate_images(latent_vectors, rotations)[0])
                for name, cfg in configs:
                    images_with_attributes[name].append(self.get_images_for_controllable_attribute(cfg, latent_vectors, rotations)[0])
                    images_without_attributes[name].append(
                        self.get_images_for_controllable_attribute(cfg, latent_vectors, rotations, other_param=True)[0])
            raw_decoded_images = np.array(raw_decoded_images)
            images_with_attributes = {k: np.array(v) for k, v in images_with_attributes.items()}
            images_without_attributes = {k: np.array(v) for k, v in images_without_attributes.items()}
        else:
            latent_vectors, rotations = self.confinet_model.encode_images(input_images)
            raw_decoded_images = self.confinet_model.generate_images(latent_vectors, rotations)
            images_with_attributes, images_without_attributes = {}, {}
            for name, cfg in configs:
                images_with_attributes[name] = self.get_images_for_controllable_attribute(cfg, latent_vectors, rotations)
                images_without_attributes[name] = self.get_images_for_controllable_attribute(cfg, latent_vectors, rotations,
                                                                                             other_param=True)
        return raw_decoded_images, images_with_attributes, images_without_attributes

    def get_metrics_for_attribute_pairs(self, set_attributes, not_set_attributes, attribute_config):
        """metrics.py:104-130 -> (mean of the driven attribute when set, when set to the other value, mean absolute
        difference of the attributes that should stay constant, correlation of the driven attribute with the setting)"""
        attribute_names = self.attribute_classifier.config["predicted_attributes"]
        driven = attribute_names.index(attribute_config.driven_attribute)
        changing = list(attribute_config.ignored_attributes) + [attribute_config.driven_attribute]
        constant = [i for i, name in enumerate(attribute_names) if name not in changing]
        mean_set = np.mean(set_attributes[:, driven])
        mean_other = np.mean(not_set_attributes[:, driven])
        n_samples = len(set_attributes)
        assert n_samples == len(not_set_attributes)
        setting = np.hstack((np.ones(n_samples), np.zeros(n_samples)))
        predicted = np.hstack((set_attributes[:, driven], not_set_attributes[:, driven]))
        corr_coef = np.corrcoef(np.vstack((setting, predicted)))
        mad = np.mean(np.mean(np.abs(set_attributes[:, constant] - not_set_attributes[:, constant]), axis=0))
        return float(mean_set), float(mean_other), float(mad), float(corr_coef[0, 1])

    def get_metrics_for_attribute_config(self, attribute_config, images_with_attribute, images_without_attribute):
        set_attributes = self.attribute_classifier.predict_attributes(images_with_attribute)
        not_set_attributes = self.attribute_classifier.predict_attributes(images_without_attribute)
        return self.get_metrics_for_attribute_pairs(set_attributes, not_set_attributes, attribute_config)

    def get_metrics(self, input_images, img_output_dir=None):
        """metrics.py:139-156; with ``img_output_dir`` every input, its reconstruction and the two images of every attribute
        configuration are written as PNG files under the reference's names (OpenCV, imported on demand)"""
        raw_decoded_images, images_with_attributes, images_without_attributes = self.generate_images_for_metric(input_images)
        if img_output_dir is not None:
            import cv2
            os.makedirs(img_output_dir, exist_ok=True)
            for i in range(len(input_images)):
                cv2.imwrite(os.path.join(img_output_dir, "gt_img_%04d.png" % i), np.asarray(input_images[i]))
                cv2.imwrite(os.path.join(img_output_dir, "raw_img_%04d.png" % i), raw_decoded_images[i])
                for config_name, _ in ControllabilityMetricConfigs.all_configs():
                    cv2.imwrite(os.path.join(img_output_dir, "%s_img_%04d.png" % (config_name, i)), images_with_attributes[config_name][i])
                    cv2.imwrite(os.path.join(img_output_dir, "%s_img_not_set_%04d.png" % (config_name, i)),
                                images_without_attributes[config_name][i])
        return self.get_metrics_from_attribute_images(images_with_attributes, images_without_attributes)

    def get_metrics_from_attribute_images(self, images_with_attributes, images_without_attributes):
        """metrics.py:158-169"""
        metrics = {}
        for name, cfg in ControllabilityMetricConfigs.all_configs():
            metrics[name] = self.get_metrics_for_attribute_config(cfg, images_with_attributes[name], images_without_attributes[name])
        metrics["contr_attribute_means"] = tuple(np.mean(list(metrics.values()), axis=0))
        metrics["controllability"] = 10 * metrics["contr_attribute_means"][2] + (1 - metrics["contr_attribute_means"][0])
        return metrics

    def update_and_log_metrics(self, images, metrics_dict, output_dir, aml_run=None, tb_log_writer=None):
        """metrics.py:171-199: appends to metrics_dict and rewrites controllability_metrics.json"""
        os.makedirs(output_dir, exist_ok=True)
        new_metrics = self.get_metrics(images)
        for key, value in new_metrics.items():
            metrics_dict.setdefault(key, []).append(value)
        if aml_run is not None:
            for key, value in new_metrics.items():
                aml_run.log(key, value)
        contr_only = dict((key, metrics_dict[key]) for key in new_metrics.keys())
        with open(os.path.join(output_dir, "controllability_metrics.json"), "w") as fp:
            json.dump(contr_only, fp, indent=4)


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
