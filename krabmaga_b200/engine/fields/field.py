"""`Field` trait (src/engine/fields/field.rs:4-9): update / lazy_update, both default no-ops."""


class Field:
    def update(self):
        pass

    def lazy_update(self):
        pass
